"""CPU restatement of the DAD-3DNet encoder ``FlameRegression.forward`` (TEST INFRASTRUCTURE ONLY; pinned to the
reference source by tests/test_oracle_pinned.py).

Follows, line by line (paths relative to /root/reference):
  model_training/model/flame_regression.py:14-106   FlameHead, FusionLayer, ClassificationHead, FlameRegression.forward
  model_training/model/encoders.py:9-59             StagedEncoder stages = [init_block, stage1, stage2, stage3, stage4]
  model_training/model/bifpn.py:11-163              BiFPN / BiFPNBlock / BiFPNDepthwiseConvBlock / BiFPNConvBlock
  model_training/config/model/resnet_regression.yaml (backbone resnet50, num_filters 256, num_classes 68, limit_value 3)
plus a restatement of the third-party backbone ``pytorchcv==0.0.65`` ``get_model("resnet50").features`` (NOT vendored;
requirements.txt:13): bottleneck ResNet-50 v1 with the stride on the FIRST 1x1 of the first unit of stages 2-4
(``conv1_stride=True``; the 3x3-stride variant is pytorchcv's ``resnet50b``), ConvBlock = Conv2d(bias=False) ->
BatchNorm2d(eps=1e-5) -> ReLU, ResInitBlock = conv7x7/2 block + MaxPool2d(3, 2, 1), ResUnit: relu(body(x) + identity(x)).

Weights come in as a ``state_dict`` with the names the reference's traced module would have
(``encoder.model.stage2.unit1.body.conv1.conv.weight`` ...), see dad_3dheads_b200/encoder_weights.py.
Everything runs in eval mode (BatchNorm uses running statistics, Dropout is the identity).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

OUTPUT_LANDMARKS_HEATMAP = "OUTPUT_LANDMARKS_HEATMAP"     # model_training/data/config.py:18
OUTPUT_3DMM_PARAMS = "OUTPUT_3DMM_PARAMS"                 # :21
OUTPUT_2D_LANDMARKS = "OUTPUT_2D_LANDMARKS"               # :16

STAGE_UNITS = (3, 4, 6, 3)
STAGE_CHANNELS = (256, 512, 1024, 2048)
BN_EPS_RESNET = 1e-5
BN_EPS_BIFPN = 4e-5          # bifpn.py:36,66
BIFPN_EPSILON = 1e-4         # bifpn.py:77
LIMIT_VALUE = 3.0            # config/model/resnet_regression.yaml:9


def _bn(x, sd, p, eps):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False,
                        0.0, eps)


def _conv_block(x, sd, p, stride=1, padding=0, act=True):
    """pytorchcv ConvBlock: conv(bias=False) -> BN(1e-5) -> ReLU."""
    x = F.conv2d(x, sd[p + ".conv.weight"], None, stride, padding)
    x = _bn(x, sd, p + ".bn", BN_EPS_RESNET)
    return F.relu(x) if act else x


def _res_unit(x, sd, p, stride, resize_identity):
    """pytorchcv ResUnit with ResBottleneck(conv1_stride=True)."""
    identity = _conv_block(x, sd, p + ".identity_conv", stride=stride, act=False) if resize_identity else x
    y = _conv_block(x, sd, p + ".body.conv1", stride=stride)
    y = _conv_block(y, sd, p + ".body.conv2", stride=1, padding=1)
    y = _conv_block(y, sd, p + ".body.conv3", act=False)
    return F.relu(y + identity)


def encoder_stages(sd, prefix="encoder.model"):
    """The five callables of StagedEncoder._get_stages (encoders.py:46-48)."""

    def init_block(x):
        x = _conv_block(x, sd, prefix + ".init_block.conv", stride=2, padding=3)
        return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)

    def make_stage(si):
        def stage(x):
            cin = 64 if si == 0 else STAGE_CHANNELS[si - 1]
            for ui in range(STAGE_UNITS[si]):
                stride = 2 if (ui == 0 and si != 0) else 1
                resize = (ui == 0)      # in != out or stride != 1
                x = _res_unit(x, sd, f"{prefix}.stage{si + 1}.unit{ui + 1}", stride, resize)
                cin = STAGE_CHANNELS[si]
            return x
        return stage

    return [init_block] + [make_stage(i) for i in range(4)]


def _dw_block(x, sd, p):
    """BiFPNDepthwiseConvBlock (bifpn.py:11-43) with the default kernel_size=1: per-channel scale -> 1x1 -> BN -> ReLU."""
    c = x.shape[1]
    x = F.conv2d(x, sd[p + ".depthwise.weight"], None, 1, 0, 1, c)
    x = F.conv2d(x, sd[p + ".pointwise.weight"], None)
    return F.relu(_bn(x, sd, p + ".bn", BN_EPS_BIFPN))


def _bifpn_block(inputs: List[torch.Tensor], sd, p):
    """BiFPNBlock.forward (bifpn.py:101-131); F.interpolate default mode is nearest."""
    p3_x, p4_x, p5_x, p6_x, p7_x = inputs
    w1 = F.relu(sd[p + ".w1"])
    w1 = w1 / torch.sum(w1, dim=0) + BIFPN_EPSILON
    w2 = F.relu(sd[p + ".w2"])
    w2 = w2 / torch.sum(w2, dim=0) + BIFPN_EPSILON
    up = lambda t, ref: F.interpolate(t, size=ref.shape[2:])
    p7_td = p7_x
    p6_td = _dw_block(w1[0, 0] * p6_x + w1[1, 0] * up(p7_td, p6_x), sd, p + ".p6_td")
    p5_td = _dw_block(w1[0, 1] * p5_x + w1[1, 1] * up(p6_td, p5_x), sd, p + ".p5_td")
    p4_td = _dw_block(w1[0, 2] * p4_x + w1[1, 2] * up(p5_td, p4_x), sd, p + ".p4_td")
    p3_td = _dw_block(w1[0, 3] * p3_x + w1[1, 3] * up(p4_td, p3_x), sd, p + ".p3_td")
    p3_out = p3_td
    p4_out = _dw_block(w2[0, 0] * p4_x + w2[1, 0] * p4_td + w2[2, 0] * up(p3_out, p4_x), sd, p + ".p4_out")
    p5_out = _dw_block(w2[0, 1] * p5_x + w2[1, 1] * p5_td + w2[2, 1] * up(p4_out, p5_x), sd, p + ".p5_out")
    p6_out = _dw_block(w2[0, 2] * p6_x + w2[1, 2] * p6_td + w2[2, 2] * up(p5_out, p6_x), sd, p + ".p6_out")
    p7_out = _dw_block(w2[0, 3] * p7_x + w2[1, 3] * p7_td + w2[2, 3] * up(p6_out, p7_x), sd, p + ".p7_out")
    return [p3_out, p4_out, p5_out, p6_out, p7_out]


def bifpn(inputs: List[torch.Tensor], sd, p="bifpn", num_layers=2):
    """BiFPN.forward (bifpn.py:152-163)."""
    c2, c3, c4 = inputs
    p3_x = F.conv2d(c2, sd[p + ".p3.weight"], sd[p + ".p3.bias"])
    p4_x = F.conv2d(c3, sd[p + ".p4.weight"], sd[p + ".p4.bias"])
    p5_x = F.conv2d(c4, sd[p + ".p5.weight"], sd[p + ".p5.bias"])
    p6_x = F.conv2d(c4, sd[p + ".p6.weight"], sd[p + ".p6.bias"], stride=2, padding=1)
    p7_x = F.conv2d(p6_x, sd[p + ".p7.conv.weight"], sd[p + ".p7.conv.bias"], stride=2, padding=1)
    p7_x = F.relu(_bn(p7_x, sd, p + ".p7.bn", BN_EPS_BIFPN))
    feats = [p3_x, p4_x, p5_x, p6_x, p7_x]
    for i in range(num_layers):
        feats = _bifpn_block(feats, sd, f"{p}.bifpn.{i}")
    return feats


def _cls_head(fmap, sd, p):
    """ClassificationHead.forward (flame_regression.py:56-59); Dropout is the identity in eval."""
    f = F.adaptive_avg_pool2d(fmap, 1).flatten(1)
    f = F.relu(F.linear(f, sd[p + ".logit_image.0.weight"], sd[p + ".logit_image.0.bias"]))
    return F.linear(f, sd[p + ".logit_image.3.weight"], sd[p + ".logit_image.3.bias"])


def flame_regression_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], return_intermediates: bool = False):
    """FlameRegression.forward (flame_regression.py:87-106).  x: [B,3,256,256] normalised image batch."""
    stages = encoder_stages(sd)
    inter = {}
    enc = []
    for i, stage in enumerate(stages[:4]):
        x = stage(x)
        enc.append(x)
        inter[f"stage{i}"] = x
    dec = bifpn(enc[1:], sd)
    for i, t in enumerate(dec):
        inter[f"p{i + 3}_out"] = t
    heatmap = F.conv2d(dec[0], sd["head.heatmap.weight"], sd["head.heatmap.bias"], padding=1)       # FlameHead :22-25
    # FusionLayer :33-42
    h, w = x.shape[2:]
    hm = F.interpolate(heatmap, size=(h, w), mode="bilinear", align_corners=True).sigmoid()
    fmap = torch.cat([x, hm, dec[2]], dim=1)
    fmap = F.conv2d(fmap, sd["fusion_layer.conv1x1.weight"], sd["fusion_layer.conv1x1.bias"])
    fmap = fmap * x
    inter["fusion"] = fmap
    fmap = stages[4](fmap)
    inter["stage4"] = fmap
    shape = _cls_head(fmap, sd, "shape").tanh() * LIMIT_VALUE
    pose = _cls_head(fmap, sd, "pose")
    lm = _cls_head(fmap, sd, "landmarks")
    B, N = lm.shape
    lm = F.relu(lm.reshape(B, N // 2, 2))
    out = {OUTPUT_LANDMARKS_HEATMAP: heatmap, OUTPUT_3DMM_PARAMS: torch.cat([shape, pose], dim=1),
           OUTPUT_2D_LANDMARKS: lm}
    return (out, inter) if return_intermediates else out
