#!/usr/bin/env python
"""Run the UNMODIFIED reference evaluator (dad_3dheads_benchmark/benchmark.py: DADEvaluator) on a ground-truth json and a
submission json and dump its (overall_result, attribute_result) -- TEST INFRASTRUCTURE (used as a subprocess: the
benchmark's own ``utils`` module name collides with the repo root's).

    python oracle/run_ref_benchmark.py <ground_truth.json> <submission.json> <out.json>

Third-party packages the image lacks come from oracle/ref_shims (fire, smplx, kaolin); without a GPU ``Tensor.cuda()`` (the
reference moves the chamfer inputs to the GPU, utils.py:139) is made the identity so the same code runs on the CPU."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as R  # noqa: E402


def main():
    gt, sub, out = [os.path.abspath(p) for p in sys.argv[1:4]]
    root = R.root()
    bdir = os.path.join(root, "dad_3dheads_benchmark")
    sys.path[:0] = [R.SHIMS, bdir]
    os.chdir(bdir)
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU evaluation (also on the GPU box: this is the CPU reference)
    import importlib.util
    src = os.path.join(bdir, "benchmark.py")
    path = src if os.path.isfile(src) else os.path.join(bdir, "benchmark.pyc")
    if path.endswith(".pyc"):
        from importlib.machinery import SourcelessFileLoader
        loader = SourcelessFileLoader("benchmark", path)
        spec = importlib.util.spec_from_loader("benchmark", loader)
    else:
        spec = importlib.util.spec_from_file_location("benchmark", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    overall, attrs = mod.DADEvaluator(gt, sub)()
    with open(out, "w") as f:
        json.dump({"overall": {k: float(v) for k, v in overall.items()},
                   "attributes": {m: {a: {str(k): float(v) for k, v in vals.items()} for a, vals in d.items()}
                                  for m, d in attrs.items()}}, f)


if __name__ == "__main__":
    main()
