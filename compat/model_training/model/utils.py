"""reference path: model_training/model/utils.py (host-side helpers of the inference path only)"""
from typing import List

import torch
import torch.nn.functional as F

from dad_3dheads_b200.predictor import calculate_paddings  # noqa: F401


def to_device(x, cuda_id: int = 0):
    return x.cuda(cuda_id) if torch.cuda.is_available() else x


def unravel_index(x: torch.Tensor) -> torch.Tensor:
    """arg-max coordinates of [B,C,H,W] heat-maps, yx order; like the reference it divides by H for both axes
    (model/utils.py:38-52), i.e. it is only right for square maps."""
    B, C, H, W = x.shape
    m = x.view(B, C, -1).argmax(-1).view(-1, 1)
    return torch.cat((torch.div(m, H, rounding_mode="trunc"), m % H), dim=1).reshape(B, C, 2)


def rot_mat_from_6dof(v: torch.Tensor) -> torch.Tensor:
    """model/utils.py:92-101 with the cross products over the last axis for every batch size."""
    assert v.shape[-1] == 6
    v = v.reshape(-1, 6)
    b1 = F.normalize(v[..., :3], dim=-1)
    b3 = F.normalize(torch.linalg.cross(b1, v[..., 3:], dim=-1), dim=-1)
    b2 = -torch.linalg.cross(b1, b3, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1)
