#!/usr/bin/env python
"""Measured accuracy of every encoder precision mode against the fp64 oracle (5 seeded images, synthetic weights), next to
the error of the reference's own fp32 CPU arithmetic (the oracle run in fp32): the number that decides which mode may call
itself fp32-class.  usage: tools/precision_report.py [--out profiles/r01_precision.md] [--layers fp16x2]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dad_3dheads_b200.encoder import Dad3dEncoder, fold_state_dict  # noqa: E402
from dad_3dheads_b200.encoder_weights import synthetic_state_dict  # noqa: E402
from oracle.encoder_oracle import (OUTPUT_2D_LANDMARKS, OUTPUT_3DMM_PARAMS, OUTPUT_LANDMARKS_HEATMAP,  # noqa: E402
                                   flame_regression_forward)

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--layers", default="", help="also print per-layer relL2 of this mode against the folded CPU executor")
ap.add_argument("--seeds", type=int, default=2, help="weight seeds to test")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def rel(x, y):
    x, y = x.double().cpu(), y.double().cpu()
    return ((x - y).norm() / y.norm()).item()


lines = ["# encoder precision modes vs the fp64 oracle (relative L2; `contract` = max over the 413 params of "
         "|err| / (1e-4 |ref| + 1e-4), must be <= 1)", "",
         "| weights seed | mode | params | landmarks | heat-map | contract | saturated |", "|---|---|---|---|---|---|---|"]
for seed in range(a.seeds):
    sd = synthetic_state_dict(seed)
    x = torch.randn(5, 3, 256, 256, generator=torch.Generator().manual_seed(42 + seed))
    with torch.no_grad():
        ref = flame_regression_forward(x.double(), {k: v.double() for k, v in sd.items()})
        ref32 = flame_regression_forward(x, sd)
    keys = (OUTPUT_3DMM_PARAMS, OUTPUT_2D_LANDMARKS, OUTPUT_LANDMARKS_HEATMAP)

    def row(name, out):
        p, r = out[OUTPUT_3DMM_PARAMS].double().cpu(), ref[OUTPUT_3DMM_PARAMS]
        contract = ((p - r).abs() / (1e-4 * r.abs() + 1e-4)).max().item()
        finite = all(torch.isfinite(out[k]).all().item() for k in keys)
        lines.append(f"| {seed} | {name} | " + " | ".join(f"{rel(out[k], ref[k]):.2e}" for k in keys) +
                     f" | {contract:.3f} | {'no' if finite else 'NON-FINITE'} |")

    row("reference fp32 on CPU (oracle in fp32)", ref32)
    for mode in ("fp32", "fp16x2", "bf16x2", "fp16", "bf16"):
        enc = Dad3dEncoder(sd, dev, precision=mode)
        row(mode, enc(x.to(dev)))
        del enc
print("\n".join(lines))

if a.layers:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.folded_ref import run_folded
    sd = synthetic_state_dict(0)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(7))
    layers, fw = fold_state_dict(sd)
    with torch.no_grad():
        ref = run_folded(x, layers, fw)
    lines += ["", f"## per-layer relL2, mode {a.layers} vs the fp32 CPU executor of the folded graph (max |activation| beside it)", "",
              "| layer | relL2 | max abs |", "|---|---|---|"]
    enc = Dad3dEncoder(sd, dev, precision=a.layers)
    enc.set_debug(True)
    enc.forward_raw(x.to(dev))
    for name in ["stem"] + [n for n, _, _ in layers if n != "stem"] + ["cat", "gap"]:
        act = enc.read_activation(name)
        r = ref[name]
        if name in ("gap", "mlp1", "mlp2"):
            got, want = act.reshape(act.shape[2], act.shape[3])[:, : r.shape[1]], r.flatten(1)
        else:
            got, want = act[..., : r.shape[1]].permute(0, 3, 1, 2), r
        lines.append(f"| {name} | {rel(got, want):.2e} | {want.abs().max().item():.3g} |")
    print("\n".join(lines[-(len(layers) + 6):]))
if a.out:
    open(os.path.join(ROOT, a.out), "w").write("\n".join(lines) + "\n")
