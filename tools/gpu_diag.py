#!/usr/bin/env python
"""First-contact GPU diagnostic for the FLAME decoder: isolates each kernel and prints error structure + timings."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dad_3dheads_b200 import HeadMesh, _lib  # noqa: E402
from oracle.flame_oracle import FlameOracle, sample_params  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def main():
    dev = torch.device("cuda", 0)
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    hm = HeadMesh()
    dec = hm.flame.decoder(dev)
    o = FlameOracle(dtype=torch.float64)
    for B in (3, 130):
        p = sample_params(B, seed=B)
        v_ref = o.vertices_3d(p)
        q_ref = o.reprojected_vertices(p)
        pd = p.to(dev)
        for mode in ("simt", "precise", "fast"):
            try:
                v3, pj = dec.decode(pd, want_vertices=True, want_projected=True, simt=(mode == "simt"),
                                    fast=(mode == "fast"))
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"B={B} {mode}: EXCEPTION {e}")
                raise
            print(f"B={B:4d} {mode:8s} vertices relL2 {rel(v3, v_ref):.3e}  projected relL2 {rel(pj, q_ref):.3e}")
            if mode != "simt" and rel(v3, v_ref) > 1e-3:
                err = (v3.double().cpu() - v_ref).abs()           # [B, V, 3]
                print("   per-row max err (first 8):", err.amax(dim=(1, 2))[:8].tolist())
                ev = err.amax(dim=(0, 2))
                blocks = ev[: (5023 // 43) * 43].reshape(-1, 43).amax(1)
                print("   per-vertex-block(43) max err (first 24):", [f"{x:.1e}" for x in blocks[:24].tolist()])
    # timing
    for B in (512, 4096, 16384):
        p = sample_params(B, seed=1).to(dev)
        for mode in ("precise", "fast"):
            for _ in range(2):
                dec.decode(p, want_vertices=True, want_projected=True, fast=(mode == "fast"))
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            n = 5
            for _ in range(n):
                dec.decode(p, want_vertices=True, want_projected=True, fast=(mode == "fast"))
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / n
            print(f"decode B={B:6d} {mode:8s}: {ms:8.3f} ms  -> {B / ms * 1e3:,.0f} heads/s")
    print("launches:", _lib.launch_count())


if __name__ == "__main__":
    main()
