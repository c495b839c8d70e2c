"""CPU restatement of the DAD-3DHeads benchmark evaluator (TEST INFRASTRUCTURE ONLY), pinned against the unmodified
reference by tests/test_evaluator_cpu.py (oracle/run_ref_benchmark.py runs dad_3dheads_benchmark/benchmark.py).

Follows dad_3dheads_benchmark/benchmark.py:17-176 (HeadAnnotation, DADEvaluator.pose_error / nme / chamfer_distance / zn /
calc_zn) and dad_3dheads_benchmark/utils.py:29-298 (get_68_landmarks, calc_ch_dist, scale_gt_to_standard, align_pred_to_gt,
procrustes).  Quirk kept as is: ``calc_zn`` sorts the gt distance matrix COLUMN-wise and then takes COLUMNS 1..n of the
index matrix (benchmark.py:124-126), so row i is compared with the i-th nearest neighbour of point j+1, not with its own
neighbours."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

SEVEN = np.array([36, 39, 42, 45, 33, 48, 54])


def landmarks68(verts: torch.Tensor, st: Dict[str, np.ndarray]) -> torch.Tensor:
    """utils.py:29-118: 17 dynamic-contour landmarks (zero pose -> row 0 of the yaw table) then 51 static ones."""
    faces = torch.from_numpy(st["faces"].astype(np.int64))

    def pts(fi, bc):
        tri = verts[faces[torch.from_numpy(fi.astype(np.int64))]]            # [L,3,3]
        return (tri * torch.from_numpy(bc).to(verts.dtype)[:, :, None]).sum(1)
    dyn = pts(st["dynamic_lmk_face_idx"][0], st["dynamic_lmk_b_coords"][0])
    sta = pts(st["static_lmk_face_idx"], st["static_lmk_b_coords"])
    return torch.cat([dyn, sta], 0)


def procrustes(X: np.ndarray, Y: np.ndarray):
    """utils.py:200-298 with scaling=True, reflection='best' -> (rotation T, scale b, translation c): Y b T + c ~ X."""
    muX, muY = X.mean(0), Y.mean(0)
    X0, Y0 = X - muX, Y - muY
    normX, normY = np.sqrt((X0 ** 2.0).sum()), np.sqrt((Y0 ** 2.0).sum())
    X0, Y0 = X0 / normX, Y0 / normY
    U, s, Vt = np.linalg.svd(X0.T @ Y0, full_matrices=False)
    T = Vt.T @ U.T
    b = s.sum() * normX / normY
    c = muX - b * (muY @ T)
    return T, b, c


def calc_zn(pred: torch.Tensor, gt: torch.Tensor, top_k: int = 5) -> float:
    """benchmark.py:110-138 for one sample [K,3] (vectorised, same index selection)."""
    d = torch.cdist(gt, gt)
    order = torch.argsort(d, dim=0)
    idx = order[:, 1:top_k + 1]                                                 # [K, top_k]
    g = gt[:, 2][:, None] >= gt[:, 2][idx]
    p = pred[:, 2][:, None] >= pred[:, 2][idx]
    return float((g == p).float().mean())


class EvaluatorOracle:
    def __init__(self, static: Dict[str, np.ndarray], head_indices: np.ndarray, face_indices: np.ndarray):
        self.st = static
        self.head = torch.from_numpy(head_indices.astype(np.int64))
        self.face = torch.from_numpy(face_indices.astype(np.int64))

    def sample(self, anno: Dict, pred: Dict) -> Dict[str, float]:
        v = np.array(anno["vertices"], dtype=np.float32)
        mv = np.array(anno["model_view_matrix"], dtype=np.float32)
        pm = np.array(anno["projection_matrix"], dtype=np.float32)
        vh = np.concatenate((v, np.ones_like(v[:, [0]])), -1)
        world = (mv @ vh.T).T                                                    # benchmark.py:45
        # pose error :80-85
        rot_180 = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]])
        R_gt = (rot_180 @ mv)[:3, :3]
        R_pred = np.array(pred["rotation_matrix"], dtype=np.float32)
        pose = float(np.linalg.norm(np.eye(3) - R_pred @ R_gt.T, "fro"))
        # nme :87-99 with landmarks_68_2d :29-37
        lm = landmarks68(torch.from_numpy(v), self.st).numpy()
        lm = np.concatenate((lm, np.ones_like(lm[:, [0]])), -1)
        lm = (pm @ (mv @ lm.T)).T
        lm = lm[:, :2] / lm[:, [3]]
        lm = np.stack((lm[:, 0], anno["image_height"] - lm[:, 1]), -1)
        p2d = np.array(pred["68_landmarks_2d"], dtype=np.float32)
        nme = float(np.mean(np.linalg.norm(lm - p2d, 2, -1) / np.sqrt(anno["bbox"][2] * anno["bbox"][3]))) * 100.0
        # chamfer :101-108 + utils.py:122-140
        gt_w = torch.from_numpy(world[:, :3].copy())
        pv = torch.tensor(pred["N_landmarks_3d"], dtype=torch.float32).view(-1, 3)
        p7 = np.array(pred["7_landmarks_3d"], dtype=np.float32).reshape(-1, 3)
        l68 = landmarks68(gt_w, self.st).numpy()
        scale = 20 / np.linalg.norm(l68[SEVEN][1] - l68[SEVEN][2])             # utils.py:166-176
        gt_s = gt_w * scale
        g7 = landmarks68(gt_s, self.st).numpy()[SEVEN]
        T, b, c = procrustes(g7, p7)
        aligned = torch.from_numpy(b * (pv.numpy().astype(np.float64) @ T) + c)
        gface = gt_s[self.face].float()
        d = torch.cdist(gface.double(), aligned.double()) ** 2
        chamfer = float(d.min(dim=1).values.float().mean())                     # w1 = 1, w2 = 0
        # z5 :140-151
        z5 = calc_zn(pv[self.head], gt_w[self.head] * -1, 5)
        return {"pose_error": pose, "nme_reprojection": nme, "z5_accuracy": z5, "chamfer": chamfer}

    def __call__(self, ground_truth: List[Dict], submission: Dict[str, Dict]) -> Dict[str, float]:
        rows = [self.sample(a, submission[a["id"]]) for a in ground_truth if a["id"] in submission]
        return {k: float(np.mean([r[k] for r in rows])) for k in rows[0]}
