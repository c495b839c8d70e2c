"""Host-side mirror of the reference's HeadMesh (model_training/head_mesh.py:9-60) over libdad3d.so."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .flame import FLAME_CONSTS, FLAMELayer, FlameParams


class HeadMesh(nn.Module):
    """Same constructor and methods as the reference.  ``decode`` is the batched single-pass entry the reference lacks
    (its predictor decodes the mesh twice per image, predictor.py:136-137)."""

    def __init__(self, flame_config: Optional[Dict[str, int]] = None, batch_size: int = 1, image_size: int = 256,
                 cuda_id: Optional[int] = None, static=None):
        super().__init__()
        self.flame_constants = FLAME_CONSTS if flame_config is None else flame_config
        self.flame = FLAMELayer(consts=self.flame_constants, batch_size=batch_size, cuda_id=cuda_id, static=static)
        self._image_size = image_size

    def flame_params(self, params_3dmm: Tensor) -> FlameParams:
        return FlameParams.from_3dmm(params_3dmm, self.flame_constants)

    def vertices_3d(self, params_3dmm: Tensor, zero_rotation: bool = False) -> Tensor:
        """head_mesh.py:28-31."""
        return self.flame.forward(self.flame_params(params_3dmm=params_3dmm), zero_rot=zero_rotation)

    def reprojected_vertices(self, params_3dmm: Tensor, to_2d: bool = True) -> Tensor:
        """head_mesh.py:33-46.  Returns [B, N, C].  Like the reference, zeroes translation z IN PLACE through the view
        of ``params_3dmm`` -- callers rely on it (predictor.py:137-142)."""
        flame_params = self.flame_params(params_3dmm=params_3dmm)
        packed = flame_params.packed().to(torch.float32)
        flame_params.translation[..., 2] = 0.0
        src_device = packed.device
        dec = self.flame.decoder(src_device)
        _, proj = dec.decode(packed.to(dec.device, non_blocking=True), want_vertices=False, want_projected=True,
                             to_2d=to_2d, image_size=float(self._image_size), hilo=self.flame.strict)
        return proj if src_device.type == "cuda" else proj.to(src_device)

    def decode(self, params_3dmm: Tensor, to_2d: bool = True, zero_rotation: bool = False,
               fast: bool = False, hilo: bool = False) -> Tuple[Tensor, Tensor]:
        """(vertices_3d, reprojected_vertices) from ONE decode pass; params_3dmm is not modified.  Batched entry point:
        the dedicated one-product decode kernel by default (relL2 ~1.5e-5, inside the 1e-4 contract); ``hilo=True`` selects
        the 3-product hi/lo blend (2e-7) that the reference-facing per-image methods above use."""
        packed = self.flame_params(params_3dmm).packed().to(torch.float32)
        src_device = packed.device
        dec = self.flame.decoder(src_device)
        v3, proj = dec.decode(packed.to(dec.device, non_blocking=True), want_vertices=True, want_projected=True,
                              to_2d=to_2d, zero_rot=zero_rotation, image_size=float(self._image_size), hilo=hilo)
        if src_device.type != "cuda":
            v3, proj = v3.to(src_device), proj.to(src_device)
        return v3, proj

    def decode_with_grad(self, params_3dmm: Tensor, to_2d: bool = True, zero_rotation: bool = False) -> Tuple[Tensor, Tensor]:
        """(vertices_3d, reprojected_vertices) as differentiable functions of ``params_3dmm`` [B,413] (a CUDA tensor that
        requires grad): what the reference's losses obtain from ``HeadMesh.vertices_3d`` / ``reprojected_vertices`` under
        autograd (losses/vertices_3d_loss.py:30-47, losses/reprojection_loss.py:22-46).  Unlike ``reprojected_vertices`` it does
        not zero translation z in the caller's tensor (its gradient is zero either way)."""
        from .flame import DecodeFunction
        packed = self.flame_params(params_3dmm).packed()
        dec = self.flame.decoder(packed.device)
        return DecodeFunction.apply(packed, dec, to_2d, zero_rotation, float(self._image_size))

    def adjust_3dmm_to_paddings(self, params_3dmm: Tensor, paddings: List[int]) -> Tensor:
        """head_mesh.py:48-60.  paddings = [pad_top, pad_bottom, pad_left, pad_right] (positive when enlarging)."""
        flame_params = self.flame_params(params_3dmm=params_3dmm)
        flame_params.translation = (
            flame_params.translation
            + Tensor([[paddings[2], paddings[0], 0]]).to(params_3dmm.device) * 2 / self._image_size
        )
        return flame_params.to_3dmm_tensor()
