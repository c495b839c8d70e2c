"""Closed-form identities of the FLAME decoder the oracle must satisfy (SURVEY §8c last row of the fixtures entry).  These
complement tests/test_oracle_pinned.py, which pins the oracle to the reference's own source."""
import numpy as np
import torch

from oracle.flame_oracle import (FLAME_CONSTS, FlameOracle, batch_rodrigues, rot_mat_from_6dof, sample_params,
                                 split_3dmm, synthetic_static)


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_zero_params_is_rotated_template(flame_static):
    o = FlameOracle(flame_static)
    p = torch.zeros(2, 413)
    p[:, 403:409] = torch.tensor([[1.0, 0, 0, 0, 1.0, 0], [0.3, -1.2, 0.5, 0.7, 0.1, -0.4]])
    v = o.vertices_3d(p)
    base = torch.from_numpy(flame_static["v_template"]).clone()
    base[:, 2] += 0.05
    R = rot_mat_from_6dof(p[:, 403:409])
    want = torch.einsum("bij,vj->bvi", R, base)
    assert _rel(v, want) < 1e-6
    assert torch.allclose(R[0], torch.eye(3), atol=1e-7)


def test_rotation_is_orthonormal():
    g = torch.Generator().manual_seed(1)
    R = rot_mat_from_6dof(torch.randn(64, 6, generator=g))
    eye = torch.eye(3).expand(64, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)


def test_rodrigues_zero_is_identity_and_matches_expm():
    R0 = batch_rodrigues(torch.zeros(3, 3))
    assert torch.equal(R0, torch.eye(3).expand(3, 3, 3))
    r = torch.tensor([[0.3, -0.2, 0.5]], dtype=torch.float64)
    K = torch.tensor([[0, -0.5, -0.2], [0.5, 0, -0.3], [0.2, 0.3, 0]], dtype=torch.float64)
    assert torch.allclose(batch_rodrigues(r)[0], torch.linalg.matrix_exp(K), atol=1e-7)


def test_jaw_zero_means_lbs_is_identity(flame_static):
    o = FlameOracle(flame_static, dtype=torch.float64)
    p = sample_params(3, seed=3).double()
    p[:, 400:403] = 0
    v = o.vertices_3d(p, zero_rotation=True)
    S = torch.from_numpy(flame_static["shapedirs"]).double()
    want = torch.from_numpy(flame_static["v_template"]).double()[None] + torch.einsum("bl,mkl->bmk", p[:, :400], S)
    want[:, :, 2] += 0.05
    assert _rel(v, want) < 1e-7          # fp32-stored skinning weights sum to 1 only to ~6e-8


def test_single_joint_closed_form(flame_static):
    """neck = eyeballs = 0: verts = (1-w_jaw) v_p + w_jaw (R_jaw (v_p - j_jaw) + j_jaw)   (SURVEY App. A)."""
    o = FlameOracle(flame_static, dtype=torch.float64)
    p = sample_params(4, seed=5).double()
    v = o.vertices_3d(p, zero_rotation=True)
    S = o.shapedirs
    v_shaped = o.v_template[None] + torch.einsum("bl,mkl->bmk", p[:, :400], S)
    J = torch.einsum("bik,ji->bjk", v_shaped, o.J_regressor)
    R = batch_rodrigues(p[:, 400:403])
    pf = torch.zeros(4, 36, dtype=torch.float64)
    pf[:, 9:18] = (R - torch.eye(3, dtype=torch.float64)).reshape(4, 9)
    v_p = v_shaped + (pf @ o.posedirs).reshape(4, -1, 3)
    w = o.lbs_weights[:, 2][None, :, None]
    jj = J[:, 2][:, None]
    want = (1 - w) * v_p + w * (torch.einsum("bij,bvj->bvi", R, v_p - jj) + jj)
    want[:, :, 2] += 0.05
    assert _rel(v, want) < 1e-7          # (1 - w_jaw) vs sum of the other fp32 weights


def test_fp32_oracle_close_to_fp64(flame_static):
    p = sample_params(8, seed=7)
    v32 = FlameOracle(flame_static).vertices_3d(p)
    v64 = FlameOracle(flame_static, dtype=torch.float64).vertices_3d(p)
    assert _rel(v32.double(), v64) < 2e-6
    q32 = FlameOracle(flame_static).reprojected_vertices(p)
    q64 = FlameOracle(flame_static, dtype=torch.float64).reprojected_vertices(p)
    assert _rel(q32.double(), q64) < 2e-6


def test_projection_equivariance(flame_static):
    o = FlameOracle(flame_static, dtype=torch.float64)
    p = sample_params(2, seed=11).double()
    q = o.reprojected_vertices(p, to_2d=False)
    p2 = p.clone()
    p2[:, 409] += 0.25                                   # translate x by 0.25 -> +32 px at image_size 256
    q2 = o.reprojected_vertices(p2, to_2d=False)
    assert torch.allclose(q2[..., 0] - q[..., 0], torch.full_like(q[..., 0], 32.0), atol=1e-9)
    assert torch.allclose(q2[..., 1:], q[..., 1:], atol=1e-12)


def test_reprojected_mutates_translation_z_like_reference(flame_static):
    o = FlameOracle(flame_static)
    p = sample_params(2, seed=2)
    assert (p[:, 411] != 0).all()
    o.reprojected_vertices(p, mutate_input=True)
    assert (p[:, 411] == 0).all()


def test_general_layout_with_neck_and_eyeballs():
    st = synthetic_static(seed=4, n_vertices=301)
    consts = dict(FLAME_CONSTS, shape=120, expression=40, neck=3, eyeballs=6)
    o = FlameOracle(st, consts=consts, dtype=torch.float64)
    p = sample_params(3, seed=9, consts=consts).double()
    assert p.shape[1] == 120 + 40 + 3 + 6 + 6 + 3 + 3 + 1
    f = split_3dmm(p, consts)
    assert f["neck"].shape[1] == 3 and f["eyeballs"].shape[1] == 6
    v = o.vertices_3d(p)
    assert torch.isfinite(v).all() and v.shape == (3, 301, 3)
    # zero pose everywhere => skinning is the identity whatever the weights
    p0 = p.clone()
    for k in ("jaw", "neck", "eyeballs"):
        split_3dmm(p0, consts)[k].zero_()
    v0 = o.vertices_3d(p0, zero_rotation=True)
    betas = torch.cat([f["shape"], torch.zeros(3, 180, dtype=torch.float64), f["expression"],
                       torch.zeros(3, 60, dtype=torch.float64)], 1)
    want = o.v_template[None] + torch.einsum("bl,mkl->bmk", betas, o.shapedirs)
    want[:, :, 2] += 0.05
    assert _rel(v0, want) < 1e-6


def test_keypoint_sets(flame_static):
    assert flame_static["keypoints_191"].shape == (191,) and flame_static["keypoints_445"].shape == (445,)
    assert np.array_equal(flame_static["indices_2d"], flame_static["keypoints_191"])
    assert len(set(flame_static["keypoints_445"].tolist())) == 445
    assert flame_static["keypoints_445"].max() < 5023
