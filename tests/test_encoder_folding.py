"""-m "not gpu": the host-side weight folding (BN fold, BiFPN depthwise fold, fusion K layout, block-diagonal heads,
fast-normalised fusion scalars) reproduces the oracle's FlameRegression.forward when executed on the CPU."""
import torch

from dad_3dheads_b200.encoder import fold_state_dict
from dad_3dheads_b200.encoder_weights import conv_specs, synthetic_state_dict
from oracle.encoder_oracle import (OUTPUT_2D_LANDMARKS, OUTPUT_3DMM_PARAMS, OUTPUT_LANDMARKS_HEATMAP,
                                   flame_regression_forward)
from tests.folded_ref import run_folded


def test_architecture_counts():
    specs = conv_specs()
    assert len(specs) == 53                                     # ResNet-50: 1 + 16*3 + 4
    sd = synthetic_state_dict(0)
    n = sum(v.numel() for k, v in sd.items() if "running" not in k)
    assert 32.5e6 < n < 33.5e6                                  # ~32.9 M parameters


def test_folded_graph_matches_oracle_fp64():
    sd = synthetic_state_dict(1)
    layers, fw = fold_state_dict(sd)
    names = [n for n, _, _ in layers]
    # 4 shortcuts fused into c3, P3 lateral composed into b0_p3td, + 8 low-res halves of the top-down nodes
    assert len(names) == len(set(names)) == 53 - 4 + 4 + 16 + 8 + 4
    x = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = flame_regression_forward(x.double(), {k: v.double() for k, v in sd.items()})
        got = run_folded(x, layers, fw)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(got["params"], ref[OUTPUT_3DMM_PARAMS]) < 2e-6    # folded weights are rounded to fp32 once
    assert rel(got["landmarks"], ref[OUTPUT_2D_LANDMARKS]) < 2e-6
    assert rel(got["heat"][:, :68], ref[OUTPUT_LANDMARKS_HEATMAP]) < 2e-6
