#!/usr/bin/env python
"""One decode pass of N heads (after a warm-up pass) -- the command profiled with ncu for the decode kernel.
    python tools/run_decode_once.py [heads] [pair]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dad_3dheads_b200 import HeadMesh  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
pair = len(sys.argv) > 2 and sys.argv[2] == "pair"
proj = len(sys.argv) > 3 and sys.argv[3] == "proj"
dev = torch.device("cuda", 0)
hm = HeadMesh(cuda_id=0)
dec = hm.flame.decoder(dev)
g = torch.Generator().manual_seed(0)
p = torch.randn(n, 413, generator=g).clamp_(-3, 3).to(dev)
p[:, 400:403] *= 0.15
for _ in range(2):
    dec.decode(p, want_vertices=True, want_projected=proj, pair=pair)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dec.decode(p, want_vertices=True, want_projected=proj, pair=pair)
e1.record()
torch.cuda.synchronize()
print(f"{n} heads, pair={pair}, proj={proj}: {e0.elapsed_time(e1):.3f} ms -> {n / e0.elapsed_time(e1) / 1e3:.2f} M heads/s")
