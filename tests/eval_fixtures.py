"""Synthetic ground-truth / submission pairs in the DAD-3DHeads benchmark's json formats (dad_3dheads_benchmark/README.md:78-90,
benchmark.py:39-61) for the evaluator tests: FLAME meshes from seeded parameters, a random rigid model-view matrix, a pinhole
projection, and a submission that is the ground truth plus seeded noise."""
import numpy as np
import torch

from oracle.evaluator_oracle import SEVEN, landmarks68
from oracle.flame_oracle import FlameOracle, load_static, sample_params


def make_pairs(n: int, seed: int = 0):
    st = load_static()
    g = np.random.default_rng(seed)
    fo = FlameOracle(st)
    p = sample_params(n, seed=seed + 100)
    p[:, 403:409] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])
    verts = fo.vertices_3d(p, zero_rotation=True).numpy()                 # model space, metres
    gts, sub = [], {}
    for i in range(n):
        a = g.normal(size=3) * 0.3
        th = np.linalg.norm(a)
        k = a / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rm = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
        mv = np.eye(4, dtype=np.float32)
        mv[:3, :3] = Rm @ np.diag([1, -1, -1])                            # camera looks down -z
        mv[:3, 3] = [g.normal() * 0.05, g.normal() * 0.05, -0.8 + g.normal() * 0.05]
        f, H, W = 1200.0, 720, 960
        pm = np.array([[f, 0, W / 2, 0], [0, f, H / 2, 0], [0, 0, 1, 0], [0, 0, 1, 0]], dtype=np.float32)
        pm[:, 2] *= -1                                                     # so that w = -z_cam > 0
        gid = f"item{i:03d}"
        gts.append({"id": gid, "vertices": verts[i].tolist(), "model_view_matrix": mv.tolist(),
                    "projection_matrix": pm.tolist(), "bbox": [300, 200, 260 + 5 * i, 300], "image_height": H,
                    "attributes": {"pose": "front" if i % 2 == 0 else "side", "occlusions": bool(i % 3 == 0)}})
        vh = np.concatenate((verts[i], np.ones((5023, 1), np.float32)), -1)
        world = (mv @ vh.T).T[:, :3]
        pred_v = (world * -1 + g.normal(size=world.shape).astype(np.float32) * 2e-3) * 1.7 + np.float32([0.1, -0.2, 0.3])
        l68 = landmarks68(torch.from_numpy(verts[i]), st).numpy()
        l68h = np.concatenate((l68, np.ones((68, 1), np.float32)), -1)
        q = (pm @ (mv @ l68h.T)).T
        q = q[:, :2] / q[:, [3]]
        lm2d = np.stack((q[:, 0], H - q[:, 1]), -1) + g.normal(size=(68, 2)) * 2.0
        p68 = landmarks68(torch.from_numpy(pred_v.astype(np.float32)), st).numpy()
        rot_180 = np.diag([1.0, -1.0, -1.0])
        R_gt = rot_180 @ mv[:3, :3]
        dR = np.eye(3) + 0.05 * K
        sub[gid] = {"68_landmarks_2d": lm2d.tolist(), "N_landmarks_3d": pred_v.astype(np.float32).tolist(),
                    "7_landmarks_3d": p68[SEVEN].tolist(), "rotation_matrix": (dR @ R_gt).tolist()}
    return gts, sub
