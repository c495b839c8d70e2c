"""-m "not gpu": the committed golden fixtures agree with the oracle that generated them (fp32 arithmetic this time)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def test_flame_golden_vs_fp32_oracle(flame_static):
    from oracle.flame_oracle import FlameOracle
    z = np.load(os.path.join(GOLDEN, "flame_decode_golden.npz"))
    o = FlameOracle(flame_static)
    p = torch.from_numpy(z["params"])
    assert _rel(o.vertices_3d(p), z["vertices3d"]) < 2e-6
    q = o.reprojected_vertices(p)
    assert _rel(q, z["projected"]) < 2e-6
    assert _rel(q[:, torch.from_numpy(z["idx445"].astype(np.int64))], z["landmarks445"]) < 2e-6


def test_encoder_golden_vs_fp32_oracle():
    from dad_3dheads_b200.encoder_weights import synthetic_state_dict
    from oracle.encoder_oracle import OUTPUT_2D_LANDMARKS, OUTPUT_3DMM_PARAMS, flame_regression_forward
    z = np.load(os.path.join(GOLDEN, "encoder_golden.npz"))
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(int(z["image_seed"])))
    with torch.no_grad():
        out = flame_regression_forward(x, synthetic_state_dict(int(z["weight_seed"])))
    assert _rel(out[OUTPUT_3DMM_PARAMS], z["params"]) < 5e-6
    assert _rel(out[OUTPUT_2D_LANDMARKS], z["landmarks"]) < 5e-6
