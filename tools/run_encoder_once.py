#!/usr/bin/env python
"""Run the encoder a few times on one synthetic batch (for `ncu -k regex:tile_gemm -s <n> -c 1` captures of one layer)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dad_3dheads_b200.encoder import Dad3dEncoder  # noqa: E402
from dad_3dheads_b200.encoder_weights import synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
dev = torch.device("cuda", 0)
enc = Dad3dEncoder(synthetic_state_dict(0), dev, precision=a.precision, want_heatmap=False)
x = torch.randn(a.batch, 3, 256, 256, device=dev)
for _ in range(a.reps):
    enc.forward_raw(x)
torch.cuda.synchronize()
print("done")
