#!/usr/bin/env python
"""Generate tests/golden/reference_*.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/ref_harness.py) in this container.

These are reference outputs, not oracle outputs: /root/reference's own predictor.py / head_mesh.py / flame.py /
model/utils.py / flame_regression.py / bifpn.py / encoders.py executed on the CPU, with shims only for the third-party
packages that are absent (oracle/ref_shims/README.md).  They pin the oracle (tests/test_oracle_pinned.py) and, on the GPU
box where /root/reference does not exist, the CUDA path (tests -m gpu).

  reference_flame.npz      HeadMesh.vertices_3d / reprojected_vertices, fp32 and fp64, B in {1, 3 (torch.cross quirk), 6}
  reference_encoder.npz    FlameRegression.forward, seed-0 synthetic weights, 2 seeded images, fp32 and fp64
  reference_predictor.npz  FaceMeshPredictor.__call__ on images/demo_heads/1.jpeg through a traced .trcd (batch-1 trace)
  reference_assets.npz     sha256 of every FLAMELayer buffer (flame.pkl -> fp32) + the landmark index sets
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402
from oracle.flame_oracle import sample_params  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", message="Using torch.cross")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def flame_golden():
    out = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        hm = R.head_mesh(dtype=dt if dt is torch.float64 else None)
        for B, seed in ((1, 11), (3, 12), (6, 13)):
            p = sample_params(B, seed=seed)
            if B == 6:
                p[0] = 0
                p[0, 403:409] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])
            out[f"params_b{B}"] = p.numpy().copy()
            q = p.clone().to(dt)
            v = hm.vertices_3d(q)                           # head_mesh.py:28-31
            vz = hm.vertices_3d(q, zero_rotation=True)
            q2 = q.clone()
            pr = hm.reprojected_vertices(q2, to_2d=False)   # head_mesh.py:33-46 (zeroes tz through the view)
            out[f"vertices3d_{name}_b{B}"] = v.to(torch.float64).numpy() if dt is torch.float64 else v.numpy()
            out[f"vertices3d_zero_rot_{name}_b{B}"] = vz.numpy()
            out[f"projected3_{name}_b{B}"] = pr.numpy()
            out[f"params_after_reproject_{name}_b{B}"] = q2.numpy()
    # keep the file small: fp64 arrays only for B=1 and B=6 vertices, stored as float64; everything else float32
    keep = {}
    for k, v in out.items():
        if "_f64_" in k and not (k.startswith("vertices3d_f64") or k.startswith("projected3_f64")):
            continue
        keep[k] = v
    np.savez_compressed(os.path.join(GOLD, "reference_flame.npz"), **keep)
    print("reference_flame.npz", os.path.getsize(os.path.join(GOLD, "reference_flame.npz")))


def encoder_golden():
    from dad_3dheads_b200.encoder_weights import synthetic_state_dict
    sd = synthetic_state_dict(0)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(777))
    out = {"image_seed": np.int64(777), "weight_seed": np.int64(0)}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = R.flame_regression(sd, dtype=dt)
        with torch.no_grad():
            o = m(x.to(dt))
        out[f"params_{name}"] = o["OUTPUT_3DMM_PARAMS"].numpy()
        out[f"landmarks_{name}"] = o["OUTPUT_2D_LANDMARKS"].numpy()
        hm = o["OUTPUT_LANDMARKS_HEATMAP"]
        out[f"heatmap_sum_{name}"] = hm.sum(dim=(2, 3)).numpy()
        out[f"heatmap_corner_{name}"] = hm[:, :, :4, :4].numpy()
    np.savez_compressed(os.path.join(GOLD, "reference_encoder.npz"), **out)
    print("reference_encoder.npz", os.path.getsize(os.path.join(GOLD, "reference_encoder.npz")))


def predictor_golden():
    import cv2
    from dad_3dheads_b200.encoder_weights import synthetic_state_dict
    pred = R.predictor(synthetic_state_dict(0))
    img = cv2.cvtColor(cv2.imread(os.path.join(GOLD, "demo_head_1.jpeg"), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    res = pred(img)
    cache = {}
    x = pred.preprocess(img, cache)
    np.savez_compressed(os.path.join(GOLD, "reference_predictor.npz"), weight_seed=np.int64(0),
                        input_sha256=sha(img), network_input_sha256=sha(x.numpy()),
                        points=res["points"], projected_vertices=res["projected_vertices"].numpy(),
                        vertices_3d=res["3d_vertices"].numpy(), params_3dmm=res["3dmm_params"].numpy())
    print("reference_predictor.npz", os.path.getsize(os.path.join(GOLD, "reference_predictor.npz")),
          {k: tuple(v.shape) for k, v in res.items()})


def assets_golden():
    buf = R.flame_buffers()
    out = {k + "_sha256": sha(v) for k, v in buf.items()}
    out.update({k + "_shape": np.asarray(v.shape) for k, v in buf.items()})
    R.activate()
    from model_training.utils import get_list_of_npy_files, load_indices_from_npy   # model_training/utils.py:81-105
    base = os.path.join(R.root(), "model_training", "model", "static", "face_keypoints")
    for sub, excl in (("191", None), ("445", "cheeks"), ("445", None)):
        files = get_list_of_npy_files({"2d_subset_path": os.path.join(base, f"keypoints_{sub}"), "2d_keys_exclude": excl})
        idx = []
        for f in sorted(files):
            idx += load_indices_from_npy(f)
        out["keypoints_" + (sub if excl or sub == "191" else "565")] = np.asarray(idx, dtype=np.int32)
    np.savez_compressed(os.path.join(GOLD, "reference_assets.npz"), **out)
    print("reference_assets.npz", os.path.getsize(os.path.join(GOLD, "reference_assets.npz")))


if __name__ == "__main__":
    assert R.available(), "needs /root/reference or oracle/_ref"
    print("reference:", R.root(), R.kind())
    flame_golden()
    encoder_golden()
    predictor_golden()
    assets_golden()
