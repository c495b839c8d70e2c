"""reference path: model_training/model/utils.py (host-side helpers of the inference path only)"""
from typing import List

import torch
import torch.nn.functional as F

from dad_3dheads_b200.predictor import calculate_paddings  # noqa: F401


def get_flame_model(flame_path=None):
    """model/utils.py:84-89: the FLAME constants as an attribute bag (``v_template, shapedirs, posedirs, J_regressor,
    kintree_table, weights, f``); a ``flame.pkl`` path is read with the package's restricted unpickler."""
    from types import SimpleNamespace

    import numpy as np

    from dad_3dheads_b200.flame import load_flame_static
    st = load_flame_static(flame_path)
    nv = st["v_template"].shape[0]
    kintree = np.stack([np.where(st["parents"] < 0, 4294967295, st["parents"]).astype(np.int64), np.arange(st["parents"].size)])
    return SimpleNamespace(v_template=st["v_template"], shapedirs=st["shapedirs"],
                           posedirs=st["posedirs"].T.reshape(nv, 3, -1), J_regressor=st["J_regressor"],
                           kintree_table=kintree, weights=st["lbs_weights"], f=st["faces"])


def get_flame_indices(name: str = "head_indices"):
    """model/utils.py:80-81: static/<name>.npy; served from the packed asset (``indices_2d``, ``flame_indices_*``)."""
    from dad_3dheads_b200.flame import load_flame_static
    st = load_flame_static()
    for key in (name, "flame_indices_" + name, "flame_indices_" + name.replace("_indices", "")):
        if key in st:
            return st[key]
    raise FileNotFoundError(f"static/{name}.npy is not part of the packed FLAME asset")


def normalize_to_cube(v: torch.Tensor) -> torch.Tensor:
    """model/utils.py:55-68: vertices [B,N,3] (or [N,3]) -> the unit cube [-1,1]^3."""
    if v.ndim == 2:
        v = v[None]
    v = v - v.min(1, True)[0]
    v = v - 0.5 * v.max(1, True)[0]
    return v / v.max(-1, True)[0].max(-2, True)[0]


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
