import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests.eval_fixtures import make_pairs
from dad_3dheads_b200.evaluator import DADEvaluatorGPU
from dad_3dheads_b200.flame import load_flame_static
from oracle.evaluator_oracle import calc_zn
gts, sub = make_pairs(2, seed=2)
st = load_flame_static()
head = torch.from_numpy(st["head_indices"].astype(np.int64))
ev = DADEvaluatorGPU()
for a in gts:
    v = np.array(a["vertices"], np.float32); mv = np.array(a["model_view_matrix"], np.float32)
    world = torch.from_numpy((mv @ np.concatenate((v, np.ones((5023,1),np.float32)),-1).T).T[:, :3].copy())
    pred = torch.tensor(sub[a["id"]]["N_landmarks_3d"], dtype=torch.float32)
    gt_h, pr_h = world[head] * -1, pred[head]
    for k in (1, 2, 3, 4, 5):
        want = calc_zn(pr_h, gt_h, k)
        got = float(ev.calc_zn(pr_h[None].cuda(), gt_h[None].cuda(), k).cpu())
        print(a["id"], "top_k", k, "oracle", round(want, 5), "gpu", round(got, 5))
    d = torch.cdist(gt_h, gt_h)
    print("self distances of points 1..5:", [float(d[c, c]) for c in range(1, 6)], "min off-diagonal", float((d + torch.eye(len(d)) * 9).min()))
    order = torch.argsort(d, dim=0)
    print("rank-0 of columns 1..5:", [int(order[0, c]) for c in range(1, 6)])
