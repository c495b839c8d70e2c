"""CPU restatement of ``predictor.FaceMeshPredictor`` (TEST INFRASTRUCTURE ONLY; pinned to the reference
source by tests/test_oracle_pinned.py).

Follows /root/reference/predictor.py:78-203 with the restated encoder (oracle/encoder_oracle.py) standing in for the
TorchScript module and the restated decoder (oracle/flame_oracle.py) for HeadMesh.  albumentations==1.0.0
(requirements.txt:16, not vendored) LongestMaxSize / PadIfNeeded / Normalize are restated with cv2 + numpy from their
published semantics.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .encoder_oracle import (OUTPUT_2D_LANDMARKS, OUTPUT_3DMM_PARAMS, flame_regression_forward)
from .flame_oracle import FLAME_CONSTS, FlameOracle

YAML_CONSTS = {"shape": 300, "expression": 100, "jaw": 3, "rotation": 6, "eyeballs": 0, "neck": 0, "translation": 3,
               "scale": 1}          # dad_3dnet.yaml:4-12 (order matters for find_3dmm_idx)


def py3round(x: float) -> int:
    # albumentations.augmentations.geometric.py3round: round-half-to-even like Python 3
    if abs(round(x) - x) == 0.5:
        return int(2.0 * round(x / 2.0))
    return int(round(x))


def calculate_paddings(orig_h: int, orig_w: int) -> List[int]:
    """model_training/model/utils.py:71-77."""
    m = max(orig_h, orig_w)
    pad_top = int((m - orig_h) / 2)
    pad_bottom = m - orig_h - pad_top
    pad_left = int((m - orig_w) / 2)
    pad_right = m - orig_w - pad_left
    return [pad_top, pad_bottom, pad_left, pad_right]


def transform(x: np.ndarray, img_size: int = 256) -> np.ndarray:
    """predictor.py:195-203."""
    import cv2
    h, w = x.shape[:2]
    scale = img_size / float(max(w, h))                                   # A.LongestMaxSize
    if scale != 1.0:
        new_h, new_w = tuple(py3round(d * scale) for d in (h, w))
        x = cv2.resize(x, dsize=(new_w, new_h), interpolation=cv2.INTER_LINEAR)
    rows, cols = x.shape[:2]                                              # A.PadIfNeeded, BORDER_CONSTANT, value None
    if rows < img_size:
        top = int((img_size - rows) / 2.0)
        bottom = img_size - rows - top
    else:
        top = bottom = 0
    if cols < img_size:
        left = int((img_size - cols) / 2.0)
        right = img_size - cols - left
    else:
        left = right = 0
    x = cv2.copyMakeBorder(x, top, bottom, left, right, cv2.BORDER_CONSTANT, value=None)
    mean = np.array([0.485, 0.456, 0.406], dtype=np.float32) * 255.0      # A.Normalize
    std = np.array([0.229, 0.224, 0.225], dtype=np.float32) * 255.0
    denominator = np.reciprocal(std, dtype=np.float32)
    img = x.astype(np.float32)
    img -= mean
    img *= denominator
    return img


def find_3dmm_idx(key: str, consts: Dict[str, int]) -> int:
    idx = 0
    for k, v in consts.items():
        if k != key:
            idx += v
        else:
            break
    return idx


class PredictorOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], static=None, consts: Optional[Dict[str, int]] = None,
                 img_size: int = 256, dtype=torch.float32):
        self.sd = {k: v.to(dtype) for k, v in state_dict.items()}
        self.consts = dict(consts or YAML_CONSTS)
        self.img_size = img_size
        self.dtype = dtype
        self.flame = FlameOracle(static, consts=self.consts, dtype=dtype, image_size=img_size)

    def encode(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            return flame_regression_forward(x.to(self.dtype), self.sd)

    def predict_batch(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Batched restatement used by the parity tests of the batched GPU entry point."""
        out = self.encode(x)
        p = out[OUTPUT_3DMM_PARAMS]
        return {"3dmm_params": p, "points": out[OUTPUT_2D_LANDMARKS] * float(self.img_size),
                "3d_vertices": self.flame.vertices_3d(p), "projected_vertices": self.flame.reprojected_vertices(p)}

    def __call__(self, image: np.ndarray) -> Dict[str, Any]:
        """predictor.py:78-83 on one HxWx3 uint8 RGB image."""
        h, w = image.shape[:2]
        x = torch.from_numpy(np.expand_dims(np.transpose(transform(image, self.img_size), (2, 0, 1)), 0))
        res = self.encode(x)
        pred_3dmm = res[OUTPUT_3DMM_PARAMS].detach().clone()
        landmarks = res[OUTPUT_2D_LANDMARKS].detach().numpy() * 256.0
        scale = self.img_size / float(max(h, w))                          # _get_paddings :117-123
        new_h, new_w = tuple(py3round(d * scale) for d in (h, w))
        paddings = calculate_paddings(new_h, new_w)
        landmarks = landmarks.clip(min=0, max=self.img_size)
        landmarks = ((landmarks - np.array([[paddings[2], paddings[0]]])) / scale).astype(int)
        si = find_3dmm_idx("scale", self.consts)                          # readjust_3dmm :154-176 (in place)
        ti = find_3dmm_idx("translation", self.consts)
        new_s = (pred_3dmm[:, si:si + 1] + 1.0) / scale - 1.0
        new_t = (pred_3dmm[:, ti:ti + 3] + 1.0
                 - torch.tensor([[paddings[2], paddings[0], 0]], dtype=pred_3dmm.dtype) * 2 / self.img_size) / scale - 1.0
        pred_3dmm[:, si:si + 1] = new_s
        pred_3dmm[:, ti:ti + 3] = new_t
        vertices_3d = self.flame.vertices_3d(pred_3dmm)[0].squeeze()
        projected = self.flame.reprojected_vertices(pred_3dmm, to_2d=True, mutate_input=True)
        return {"points": np.reshape(landmarks, (-1, 2)), "projected_vertices": projected, "3d_vertices": vertices_3d,
                "3dmm_params": pred_3dmm}
