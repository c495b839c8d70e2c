#!/usr/bin/env python
"""GPU bring-up diagnostic for the encoder: per-layer comparison against the CPU folded-graph executor + timings."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dad_3dheads_b200 import _lib  # noqa: E402
from dad_3dheads_b200.encoder import Dad3dEncoder, fold_state_dict  # noqa: E402
from dad_3dheads_b200.encoder_weights import synthetic_state_dict  # noqa: E402
from tests.folded_ref import run_folded  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main():
    dev = torch.device("cuda", 0)
    sd = synthetic_state_dict(0)
    layers, fw = fold_state_dict(sd)
    B = 2
    x = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    t0 = time.time()
    with torch.no_grad():
        ref = run_folded(x, layers, fw)
    print(f"cpu folded fp64 reference: {time.time() - t0:.1f}s")
    order = [n for n, _, _ in layers]
    for prec in ("fp32", "bf16x2", "bf16"):
        enc = Dad3dEncoder(sd, dev, precision=prec)
        enc.set_debug(True)
        try:
            params, lms, heat = enc.forward_raw(x.to(dev))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[{prec}] forward EXCEPTION: {e}")
            raise
        print(f"[{prec}] params rel {rel(params, ref['params']):.3e}  landmarks rel {rel(lms, ref['landmarks']):.3e}  "
              f"heat rel {rel(heat, ref['heat'][:, :68]):.3e}")
        if prec == "fp32" or rel(params, ref["params"]) > 0.1:
            names = ["stem_conv", "stem"] + [n for n in order if n not in ("stem",)] + ["cat", "gap"]
            bad = 0
            for n in names:
                if n not in ref:
                    continue
                try:
                    a = enc.read_activation(n if n != "heat" else "heat")
                except Exception as e:  # noqa: BLE001
                    print(f"   {n:10s} read failed: {e}")
                    continue
                r = ref[n]                                       # NCHW
                if n in ("gap", "mlp1", "mlp2"):
                    got = a.reshape(a.shape[2], a.shape[3])[:, : r.shape[1]]
                    want = r.flatten(1)
                else:
                    got = a[..., : r.shape[1]].permute(0, 3, 1, 2)
                    want = r
                e = rel(got, want)
                flag = "" if e < (1e-5 if prec == "fp32" else 1e-1) else "   <-- MISMATCH"
                if flag or n in ("stem", "s1u3c3", "s2u4c3", "s3u6c3", "b1_p3td", "fusion", "s4u3c3", "mlp2"):
                    print(f"   {n:10s} {tuple(a.shape)} rel {e:.3e}{flag}")
                if flag:
                    bad += 1
                    if bad >= 6:
                        break
        del enc
    # timings
    for prec in ("fp32", "bf16x2", "bf16"):
        enc = Dad3dEncoder(sd, dev, precision=prec, want_heatmap=False)
        for B in (8, 64):
            xb = torch.randn(B, 3, 256, 256, device=dev)
            for _ in range(2):
                enc.forward_raw(xb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3
            for _ in range(n):
                enc.forward_raw(xb)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"encoder {prec:7s} B={B:3d}: {ms:8.3f} ms -> {B / ms * 1e3:8.0f} img/s  "
                  f"({15.12e9 * B / ms / 1e9:.1f} useful TFLOP/s)")
        del enc
    print("launches:", _lib.launch_count())


if __name__ == "__main__":
    main()
