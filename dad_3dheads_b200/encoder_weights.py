"""Weights of the DAD-3DNet encoder in the reference's naming (what ``torch.jit.load(dad_3dheads.trcd).state_dict()``
yields: pytorchcv ResNet-50 features + BiFPN + heads), plus a seeded synthetic initialiser.

The released checkpoint is downloaded on first use by the reference (predictor.py:21-26,205-211) and cannot be fetched
offline, so parity and benchmarks use ``synthetic_state_dict(seed)``: random-init weights of exactly that architecture,
scaled so activations stay O(1) through all ~55 layers (BatchNorm running statistics randomised too).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

STAGE_UNITS = (3, 4, 6, 3)
STAGE_CHANNELS = (256, 512, 1024, 2048)
NUM_FILTERS = 256      # config/model/resnet_regression.yaml: num_filters
NUM_CLASSES = 68       # heat-map channels / 2D landmarks
LINEAR_SIZE = 512      # ClassificationHead linear_size
HEAD_OUT = (("shape", 403), ("pose", 10), ("landmarks", 2 * NUM_CLASSES))


def conv_specs() -> List[Tuple[str, int, int, int]]:
    """(prefix, cin, cout, k) of every pytorchcv ConvBlock (conv bias=False + BN) of the backbone, in forward order."""
    specs = [("encoder.model.init_block.conv", 3, 64, 7)]
    cin = 64
    for si, (nu, cout) in enumerate(zip(STAGE_UNITS, STAGE_CHANNELS)):
        mid = cout // 4
        for ui in range(nu):
            p = f"encoder.model.stage{si + 1}.unit{ui + 1}"
            specs.append((p + ".body.conv1", cin, mid, 1))
            specs.append((p + ".body.conv2", mid, mid, 3))
            specs.append((p + ".body.conv3", mid, cout, 1))
            if ui == 0:
                specs.append((p + ".identity_conv", cin, cout, 1))
            cin = cout
    return specs


def synthetic_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def uniform(n, lo, hi):
        return torch.rand(n, generator=g) * (hi - lo) + lo

    def bn(p, c, gamma_lo=0.8, gamma_hi=1.2):
        sd[p + ".weight"] = uniform(c, gamma_lo, gamma_hi)
        sd[p + ".bias"] = randn(c, std=0.1)
        sd[p + ".running_mean"] = randn(c, std=0.1)
        sd[p + ".running_var"] = uniform(c, 0.5, 1.5)

    for p, cin, cout, k in conv_specs():
        fan_in = cin * k * k
        sd[p + ".conv.weight"] = randn(cout, cin, k, k, std=(2.0 / fan_in) ** 0.5)
        if p.endswith("conv3"):
            bn(p + ".bn", cout, 0.2, 0.5)        # damped residual branch keeps the 16-unit stack O(1)
        elif p.endswith("identity_conv"):
            bn(p + ".bn", cout, 0.5, 0.8)
        else:
            bn(p + ".bn", cout)

    f = NUM_FILTERS
    for name, cin in (("p3", 256), ("p4", 512), ("p5", 1024)):
        sd[f"bifpn.{name}.weight"] = randn(f, cin, 1, 1, std=(1.0 / cin) ** 0.5)
        sd[f"bifpn.{name}.bias"] = randn(f, std=0.1)
    sd["bifpn.p6.weight"] = randn(f, 1024, 3, 3, std=(1.0 / (1024 * 9)) ** 0.5)
    sd["bifpn.p6.bias"] = randn(f, std=0.1)
    sd["bifpn.p7.conv.weight"] = randn(f, f, 3, 3, std=(2.0 / (f * 9)) ** 0.5)
    sd["bifpn.p7.conv.bias"] = randn(f, std=0.1)
    bn("bifpn.p7.bn", f)
    for li in range(2):
        for node in ("p3_td", "p4_td", "p5_td", "p6_td", "p4_out", "p5_out", "p6_out", "p7_out"):
            p = f"bifpn.bifpn.{li}.{node}"
            sd[p + ".depthwise.weight"] = uniform(f, 0.5, 1.5).reshape(f, 1, 1, 1)
            sd[p + ".pointwise.weight"] = randn(f, f, 1, 1, std=(2.0 / f) ** 0.5)
            bn(p + ".bn", f)
        sd[f"bifpn.bifpn.{li}.w1"] = uniform(8, 0.5, 1.5).reshape(2, 4)
        sd[f"bifpn.bifpn.{li}.w2"] = uniform(12, 0.5, 1.5).reshape(3, 4)
        sd[f"bifpn.bifpn.{li}.w1"][0, 1] = -0.3      # exercises the relu() on the fusion weights (bifpn.py:105)

    sd["head.heatmap.weight"] = randn(NUM_CLASSES, f, 3, 3, std=(1.0 / (f * 9)) ** 0.5)
    sd["head.heatmap.bias"] = randn(NUM_CLASSES, std=0.1)
    cat = f + NUM_CLASSES + 1024
    sd["fusion_layer.conv1x1.weight"] = randn(1024, cat, 1, 1, std=(1.0 / cat) ** 0.5)
    sd["fusion_layer.conv1x1.bias"] = randn(1024, std=0.1) + 0.5
    for name, nout in HEAD_OUT:
        sd[f"{name}.logit_image.0.weight"] = randn(LINEAR_SIZE, 2048, std=0.1 * (2.0 / 2048) ** 0.5)
        sd[f"{name}.logit_image.0.bias"] = randn(LINEAR_SIZE, std=0.1)
        sd[f"{name}.logit_image.3.weight"] = randn(nout, LINEAR_SIZE, std=(1.0 / LINEAR_SIZE) ** 0.5)
        sd[f"{name}.logit_image.3.bias"] = randn(nout, std=0.1)
    # keep the pose head's scale/translation outputs in a sane range (translation ~ +-0.3, scale ~ +-0.5)
    sd["pose.logit_image.3.weight"] *= 0.5
    sd["landmarks.logit_image.3.weight"] *= 0.3
    sd["landmarks.logit_image.3.bias"] += 0.5
    return {k: v.contiguous().float() for k, v in sd.items()}
