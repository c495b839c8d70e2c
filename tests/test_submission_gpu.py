"""-m gpu: benchmark-submission fields against an oracle restatement of get_68_landmarks etc."""
import numpy as np
import pytest
import torch

from dad_3dheads_b200.encoder_weights import synthetic_state_dict
from oracle.flame_oracle import load_static, rot_mat_from_6dof
from oracle.predictor_oracle import PredictorOracle

pytestmark = pytest.mark.gpu


def _oracle_68(verts: torch.Tensor, st) -> torch.Tensor:
    """model_training/data/utils.py:120-206 restated: 17 dynamic (zero-pose row 0) + 51 static barycentric landmarks."""
    faces = torch.from_numpy(st["faces"].astype(np.int64))

    def bary_pts(face_idx, b):
        tri = verts[faces[face_idx]]                       # [L,3(verts),C]
        return (tri * b[:, :, None]).sum(1)

    dyn = bary_pts(torch.from_numpy(st["dynamic_lmk_face_idx"][0].astype(np.int64)),
                   torch.from_numpy(st["dynamic_lmk_b_coords"][0]).to(verts.dtype))
    sta = bary_pts(torch.from_numpy(st["static_lmk_face_idx"].astype(np.int64)),
                   torch.from_numpy(st["static_lmk_b_coords"]).to(verts.dtype))
    return torch.cat([dyn, sta], 0)


def test_submission_fields(cuda_device, tmp_path):
    from dad_3dheads_b200.predictor import FaceMeshPredictor
    from dad_3dheads_b200.submission import SEVEN_OF_68, SubmissionWriter
    sd = synthetic_state_dict(0)
    st = load_static()
    pred = FaceMeshPredictor.dad_3dnet(state_dict=sd)
    sw = SubmissionWriter(pred)
    x = torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(11))
    sub = sw.predict(x, ["a", "b", "c"])
    assert set(sub) == {"a", "b", "c"}
    want = PredictorOracle(sd, dtype=torch.float64).predict_batch(x)
    for i, key in enumerate(["a", "b", "c"]):
        item = sub[key]
        assert set(item) == {"68_landmarks_2d", "N_landmarks_3d", "7_landmarks_3d", "rotation_matrix"}
        lm3 = _oracle_68(want["3d_vertices"][i], st)
        lm2 = _oracle_68(want["projected_vertices"][i], st)
        got3 = torch.tensor(item["7_landmarks_3d"], dtype=torch.float64)
        assert got3.shape == (7, 3) and (got3 - lm3[list(SEVEN_OF_68)]).abs().max() < 1e-5
        got2 = torch.tensor(item["68_landmarks_2d"], dtype=torch.float64)
        assert got2.shape == (68, 2) and (got2 - lm2).abs().max() < 2e-2                    # pixels
        assert len(item["N_landmarks_3d"]) == 5023
        R = rot_mat_from_6dof(want["3dmm_params"][i:i + 1, 403:409])[0]
        assert (torch.tensor(item["rotation_matrix"], dtype=torch.float64) - R).abs().max() < 1e-4
    SubmissionWriter.save(sub, str(tmp_path / "sub.json"))
    import json
    assert set(json.load(open(tmp_path / "sub.json"))) == {"a", "b", "c"}


def test_submission_2d_landmarks_are_in_original_image_pixels(cuda_device):
    """Non-square, non-256 inputs: "68_landmarks_2d" must be in ORIGINAL-image pixels (what benchmark.py:86-99 compares with
    its ground truth).  Oracle route = the reference's: per image, predictor.__call__ semantics (readjust the 3DMM vector to
    the input image, then reprojected_vertices) followed by the barycentric 68 embedding on those projected vertices."""
    from dad_3dheads_b200.predictor import FaceMeshPredictor
    from dad_3dheads_b200.submission import SubmissionWriter
    sd = synthetic_state_dict(0)
    st = load_static()
    pred = FaceMeshPredictor.dad_3dnet(state_dict=sd)
    sw = SubmissionWriter(pred)
    g = np.random.default_rng(3)
    imgs = [g.integers(0, 256, s + (3,), dtype=np.uint8) for s in ((300, 517), (641, 203), (256, 256))]
    sub = sw.predict(imgs, ["a", "b", "c"])
    orc = PredictorOracle(sd, dtype=torch.float64)
    for key, img in zip(["a", "b", "c"], imgs):
        want = orc(img)                                        # projected_vertices in input-image pixels (predictor.py:137)
        lm2 = _oracle_68(want["projected_vertices"][0], st)
        got2 = torch.tensor(sub[key]["68_landmarks_2d"], dtype=torch.float64)
        assert (got2 - lm2).abs().max() < 0.1, (key, (got2 - lm2).abs().max())          # pixels of a ~500-px image
