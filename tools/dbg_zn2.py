import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests.eval_fixtures import make_pairs
from dad_3dheads_b200.evaluator import DADEvaluatorGPU
gts, sub = make_pairs(5, seed=2)
ev = DADEvaluatorGPU()
res = ev.metrics(gts, [sub[a["id"]] for a in gts])
print("batched z5", res["z5"])
for i, a in enumerate(gts):
    r = ev.metrics([a], [sub[a["id"]]])
    print(a["id"], "single z5", r["z5"], "chamfer", r["chamfer"], res["chamfer"][i])
