#!/usr/bin/env python
"""Generate tests/golden/flame_decode_golden.npz from the fp64 oracle (the reference itself cannot run here -- see
oracle/__init__.py -- so these are regression pins of the restatement, not reference outputs)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.flame_oracle import FlameOracle, load_static, sample_params  # noqa: E402


def main():
    st = load_static()
    o = FlameOracle(st, dtype=torch.float64)
    p = sample_params(6, seed=2024)
    p[0] = 0
    p[0, 403:409] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])
    v = o.vertices_3d(p)
    q = o.reprojected_vertices(p)
    idx = st["keypoints_445"]
    out = os.path.join(ROOT, "tests", "golden", "flame_decode_golden.npz")
    np.savez_compressed(out, params=p.numpy(), vertices3d=v.float().numpy(), projected=q.float().numpy(), idx445=idx,
                        landmarks445=q.float().numpy()[:, idx])
    print("wrote", out, os.path.getsize(out))


def encoder_golden():
    """Encoder golden: seed-0 synthetic weights, 2 seeded images -> params / landmarks from the fp64 oracle."""
    from dad_3dheads_b200.encoder_weights import synthetic_state_dict
    from oracle.encoder_oracle import (OUTPUT_2D_LANDMARKS, OUTPUT_3DMM_PARAMS, OUTPUT_LANDMARKS_HEATMAP,
                                       flame_regression_forward)
    sd = {k: v.double() for k, v in synthetic_state_dict(0).items()}
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(777))
    with torch.no_grad():
        out = flame_regression_forward(x.double(), sd)
    path = os.path.join(ROOT, "tests", "golden", "encoder_golden.npz")
    np.savez_compressed(path, image_seed=np.int64(777), weight_seed=np.int64(0),
                        params=out[OUTPUT_3DMM_PARAMS].float().numpy(), landmarks=out[OUTPUT_2D_LANDMARKS].float().numpy(),
                        heatmap_checksum=out[OUTPUT_LANDMARKS_HEATMAP].sum(dim=(2, 3)).float().numpy())
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
    encoder_golden()
