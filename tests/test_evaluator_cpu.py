"""-m "not gpu": the evaluator oracle (oracle/evaluator_oracle.py) against the UNMODIFIED reference evaluator
(dad_3dheads_benchmark/benchmark.py::DADEvaluator, run by oracle/run_ref_benchmark.py with shims for fire / smplx / kaolin)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_harness as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not R.available(), reason="reference tree not available")
def test_evaluator_oracle_matches_reference(tmp_path):
    from dad_3dheads_b200.flame import load_flame_static
    from oracle.evaluator_oracle import EvaluatorOracle
    from tests.eval_fixtures import make_pairs
    gts, sub = make_pairs(3, seed=1)
    json.dump(gts, open(tmp_path / "gt.json", "w"))
    json.dump(sub, open(tmp_path / "sub.json", "w"))
    out = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "oracle", "run_ref_benchmark.py"),
                          str(tmp_path / "gt.json"), str(tmp_path / "sub.json"), str(tmp_path / "ref.json")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    ref = json.load(open(tmp_path / "ref.json"))["overall"]
    st = load_flame_static()
    got = EvaluatorOracle(st, st["head_indices"], st["flame_indices_face"])(gts, sub)
    assert set(got) == set(ref)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 1e-5 * abs(ref[k]) + 1e-6, (k, got[k], ref[k])
    assert 0.5 < ref["z5_accuracy"] <= 1.0 and ref["chamfer"] > 0 and ref["nme_reprojection"] > 0
