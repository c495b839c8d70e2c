"""CPU restatement of the reference's FLAME head decoder + weak-perspective projection (TEST INFRASTRUCTURE ONLY).

Pinned against the unmodified reference source (oracle/ref_harness.py, tests/test_oracle_pinned.py).

Each function cites the reference file:line (relative to /root/reference) it follows.  ``smplx.lbs`` (smplx==0.1.26,
pinned in requirements.txt:17, NOT vendored in the reference tree) is restated from its published algorithm
(Loper et al. SMPL; the FLAME_PyTorch layer) -- call site model_training/model/flame.py:212-221.

Everything is plain torch CPU; ``dtype`` selects float32 (the reference's arithmetic) or float64 (error yard-stick).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

# model_training/model/flame.py:17-26 (dict order there is rotation-before-jaw; slicing order below is what matters)
FLAME_CONSTS = {"shape": 300, "expression": 100, "rotation": 6, "jaw": 3, "eyeballs": 0, "neck": 0,
                "translation": 3, "scale": 1}
MESH_OFFSET_Z = 0.05          # flame.py:114
MAX_SHAPE, MAX_EXPRESSION = 300, 100   # flame.py:107-108

_DEFAULT_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dad_3dheads_b200", "assets",
                              "flame_static.npz")


def load_static(path: Optional[str] = None) -> Dict[str, np.ndarray]:
    """The packed fp32 FLAME constants (tools/pack_flame_assets.py; what flame.py:124-180 registers as buffers)."""
    with np.load(path or _DEFAULT_ASSET) as z:
        return {k: z[k] for k in z.files}


def synthetic_static(seed: int = 0, n_vertices: int = 5023, n_shape: int = 400, n_joints: int = 5) -> Dict[str, np.ndarray]:
    """A seeded stand-in with the FLAME shapes, for tests that must not depend on the real asset."""
    g = np.random.default_rng(seed)
    v = (g.standard_normal((n_vertices, 3)) * 0.08).astype(np.float32)
    sd = (g.standard_normal((n_vertices, 3, n_shape)) * 2e-3 * np.linspace(1.0, 0.05, n_shape)).astype(np.float32)
    pd = (g.standard_normal(((n_joints - 1) * 9, n_vertices * 3)) * 1e-3).astype(np.float32)
    jr = np.zeros((n_joints, n_vertices), np.float32)
    for j in range(n_joints):
        idx = g.choice(n_vertices, 9, replace=False)
        w = g.random(9).astype(np.float32)
        jr[j, idx] = w / w.sum()
    w = np.zeros((n_vertices, n_joints), np.float32)
    for i in range(n_vertices):
        k = g.integers(1, 4)
        idx = g.choice(n_joints, k, replace=False)
        ww = g.random(k).astype(np.float32)
        w[i, idx] = ww / ww.sum()
    return dict(v_template=v, shapedirs=sd, posedirs=pd, J_regressor=jr, parents=np.array([-1, 0, 1, 1, 1][:n_joints], np.int32),
                lbs_weights=w)


# ------------------------------------------------------------------------------------------------------------------
# model_training/model/utils.py:92-101
def rot_mat_from_6dof(v: torch.Tensor) -> torch.Tensor:
    """6-DoF -> rotation matrix by Gram-Schmidt; b1,b2,b3 are the COLUMNS (stack on dim=-1).

    Quirk: the reference calls ``torch.cross`` with no ``dim`` which picks the first size-3 axis, so
    it is wrong for batch == 3.  The oracle uses the evidently intended last axis for every batch size.
    """
    assert v.shape[-1] == 6
    v = v.reshape(-1, 6)
    vx, vy = v[:, :3], v[:, 3:]
    b1 = torch.nn.functional.normalize(vx, dim=-1)                       # eps 1e-12 (F.normalize default)
    b3 = torch.nn.functional.normalize(torch.linalg.cross(b1, vy, dim=-1), dim=-1)
    b2 = -torch.linalg.cross(b1, b3, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1)


# smplx/lbs.py batch_rodrigues (0.1.26)
def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """Axis-angle [N,3] -> [N,3,3].  The 1e-8 is added to the VECTOR inside the norm (App. D.9)."""
    angle = torch.linalg.norm(rot_vecs + 1e-8, dim=1, keepdim=True)      # [N,1]
    d = rot_vecs / angle
    s, c = torch.sin(angle)[:, :, None], torch.cos(angle)[:, :, None]    # [N,1,1]
    z = torch.zeros_like(d[:, 0])
    K = torch.stack([z, -d[:, 2], d[:, 1], d[:, 2], z, -d[:, 0], -d[:, 1], d[:, 0], z], dim=1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return eye + s * K + (1.0 - c) * torch.bmm(K, K)


# smplx/lbs.py batch_rigid_transform (0.1.26)
def batch_rigid_transform(rot_mats: torch.Tensor, joints: torch.Tensor, parents) -> torch.Tensor:
    """rot_mats [B,J,3,3], joints [B,J,3] -> relative transforms A [B,J,4,4] (rest joint removed from the translation)."""
    B, J = joints.shape[:2]
    rel = joints.clone()
    for i in range(1, J):
        rel[:, i] = joints[:, i] - joints[:, int(parents[i])]
    M = torch.zeros(B, J, 4, 4, dtype=joints.dtype)
    M[:, :, :3, :3] = rot_mats
    M[:, :, :3, 3] = rel
    M[:, :, 3, 3] = 1.0
    chain = [M[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], M[:, i]))
    G = torch.stack(chain, dim=1)                                        # [B,J,4,4]
    jh = torch.cat([joints, torch.zeros(B, J, 1, dtype=joints.dtype)], dim=2)[..., None]   # [B,J,4,1]
    corr = torch.matmul(G, jh)                                           # [B,J,4,1]
    A = G.clone()
    A[:, :, :, 3:4] = A[:, :, :, 3:4] - corr
    return A


# smplx/lbs.py lbs (0.1.26), call site flame.py:212-221
def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    B = betas.shape[0]
    v_shaped = v_template[None] + torch.einsum("bl,mkl->bmk", betas, shapedirs)          # blend_shapes
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)                               # vertices2joints
    R = batch_rodrigues(pose.reshape(-1, 3)).reshape(B, -1, 3, 3)
    pose_feature = (R[:, 1:] - torch.eye(3, dtype=betas.dtype)).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).reshape(B, -1, 3)
    A = batch_rigid_transform(R, J, parents)
    nj = J_regressor.shape[0]
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.reshape(B, nj, 16)).reshape(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype)], dim=2)
    out = torch.matmul(T, vh[..., None])[:, :, :3, 0]
    return out


# model_training/model/flame.py:41-84
def split_3dmm(p: torch.Tensor, consts: Dict[str, int]) -> Dict[str, torch.Tensor]:
    """VIEWS into the [B,413] tensor in the hard-coded order shape, expression, jaw, rotation, eyeballs, neck,
    translation, scale (NOT the dict order)."""
    assert p.ndim == 2
    out, cur = {}, 0
    for k in ("shape", "expression", "jaw", "rotation", "eyeballs", "neck", "translation", "scale"):
        out[k] = p[:, cur:cur + consts[k]]
        cur += consts[k]
    return out


class FlameOracle:
    """FLAMELayer.forward (flame.py:182-229) + HeadMesh.{vertices_3d,reprojected_vertices} (head_mesh.py:28-46)."""

    def __init__(self, static: Optional[Dict[str, np.ndarray]] = None, consts: Optional[Dict[str, int]] = None,
                 dtype=torch.float32, image_size: int = 256):
        st = static if static is not None else load_static()
        self.consts = dict(consts or FLAME_CONSTS)
        self.dtype = dtype
        self.image_size = image_size
        t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
        self.v_template = t(st["v_template"])
        self.shapedirs = t(st["shapedirs"])
        self.posedirs = t(st["posedirs"])
        self.J_regressor = t(st["J_regressor"])
        self.parents = [int(x) for x in st["parents"]]
        self.lbs_weights = t(st["lbs_weights"])

    def flame_forward(self, fp: Dict[str, torch.Tensor], zero_rot=False, zero_jaw=False) -> torch.Tensor:
        B = fp["shape"].shape[0]
        z = lambda n: torch.zeros(B, n, dtype=self.dtype)
        # flame.py:191-200 -- remaining betas are registered zeros
        betas = torch.cat([fp["shape"], z(MAX_SHAPE - self.consts["shape"]),
                           fp["expression"], z(MAX_EXPRESSION - self.consts["expression"])], dim=1)
        neck = fp["neck"] if fp["neck"].shape[1] else z(3)                # flame.py:201-203
        eyes = fp["eyeballs"] if fp["eyeballs"].shape[1] else z(6)
        jaw = fp["jaw"] if fp["jaw"].shape[1] else z(3)
        if zero_jaw:
            jaw = torch.zeros_like(jaw)
        full_pose = torch.cat([z(3), neck, jaw, eyes], dim=1)             # flame.py:205-208 (global rot NOT given to lbs)
        v = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs, self.J_regressor, self.parents,
                self.lbs_weights)
        v = v.clone()
        v[:, :, 2] += MESH_OFFSET_Z                                       # flame.py:224
        if not zero_rot:                                                  # flame.py:225-228
            R = rot_mat_from_6dof(fp["rotation"]).to(v.dtype)
            v = torch.matmul(R[:, None], v[..., None])[..., 0]
        return v

    def vertices_3d(self, params: torch.Tensor, zero_rotation=False) -> torch.Tensor:
        """head_mesh.py:28-31."""
        return self.flame_forward(split_3dmm(params.to(self.dtype), self.consts), zero_rot=zero_rotation)

    def reprojected_vertices(self, params: torch.Tensor, to_2d=True, mutate_input=False) -> torch.Tensor:
        """head_mesh.py:33-46.  The reference zeroes translation z THROUGH the view (App. D.1); ``mutate_input``
        reproduces that side effect on ``params`` (only meaningful when params already has self.dtype)."""
        p = params if (mutate_input and params.dtype == self.dtype) else params.to(self.dtype).clone()
        fp = split_3dmm(p, self.consts)
        v = self.flame_forward(fp, zero_rot=False)
        scale = torch.clamp(fp["scale"][:, None] + 1.0, min=1e-8)
        v = v * scale
        fp["translation"][..., 2] = 0.0
        v = v + fp["translation"][:, None]
        proj = (v + 1.0) / 2.0 * self.image_size
        return proj[..., :2] if to_2d else proj

    @staticmethod
    def gather_landmarks(projected: torch.Tensor, idx) -> torch.Tensor:
        """demo_utils.py:37-47 (np.take along the vertex axis; the int truncation there is drawing-only)."""
        return projected[:, torch.as_tensor(np.asarray(idx), dtype=torch.long)]


def sample_params(n: int, seed: int = 0, consts: Optional[Dict[str, int]] = None) -> torch.Tensor:
    """SURVEY §8(d) config-5 distribution of 413-vectors: shape/expr ~N(0,1) clipped to +-3, jaw ~N(0,0.15) clipped,
    rotation 6-vector ~N(0,1), translation ~U(-0.3,0.3), scale ~U(-0.5,0.5); eyeballs/neck (if any) ~N(0,0.1)."""
    c = dict(consts or FLAME_CONSTS)
    g = torch.Generator().manual_seed(seed)
    parts = []
    for k in ("shape", "expression", "jaw", "rotation", "eyeballs", "neck", "translation", "scale"):
        w = c[k]
        if k in ("shape", "expression"):
            x = torch.randn(n, w, generator=g).clamp_(-3, 3)
        elif k == "jaw":
            x = (torch.randn(n, w, generator=g) * 0.15).clamp_(-3, 3)
        elif k == "rotation":
            x = torch.randn(n, w, generator=g)
        elif k in ("eyeballs", "neck"):
            x = torch.randn(n, w, generator=g) * 0.1
        elif k == "translation":
            x = torch.rand(n, w, generator=g) * 0.6 - 0.3
        else:
            x = torch.rand(n, w, generator=g) - 0.5
        parts.append(x)
    return torch.cat(parts, dim=1).contiguous()
