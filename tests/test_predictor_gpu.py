"""-m gpu: the reference-facing API (FaceMeshPredictor) end to end against the oracle restatement of predictor.py."""
import os

import numpy as np
import pytest
import torch

from dad_3dheads_b200.encoder_weights import synthetic_state_dict
from oracle.predictor_oracle import PredictorOracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def predictor(cuda_device):
    from dad_3dheads_b200.predictor import FaceMeshPredictor
    return FaceMeshPredictor.dad_3dnet(state_dict=synthetic_state_dict(0))


@pytest.fixture(scope="module")
def oracle():
    return PredictorOracle(synthetic_state_dict(0), dtype=torch.float64)


def _demo_image():
    import cv2
    return cv2.cvtColor(cv2.imread(os.path.join(GOLDEN, "demo_head_1.jpeg")), cv2.COLOR_BGR2RGB)


def test_single_image_call_matches_reference_semantics(predictor, oracle):
    """BASELINE configs[0]: FaceMeshPredictor.__call__ on images/demo_heads/1.jpeg (954x766 -> letter-boxed 256x206)."""
    img = _demo_image()
    got = predictor(img)
    want = oracle(img)
    assert set(got) == {"points", "projected_vertices", "3d_vertices", "3dmm_params"}
    assert got["points"].shape == (68, 2) and got["points"].dtype.kind == "i"
    assert got["projected_vertices"].shape == (1, 5023, 2) and got["projected_vertices"].device.type == "cpu"
    assert got["3d_vertices"].shape == (5023, 3) and got["3dmm_params"].shape == (1, 413)
    assert got["3dmm_params"][0, 411].item() == 0.0                      # tz zeroed in place (head_mesh.py:41)
    assert _rel(got["3dmm_params"], want["3dmm_params"]) < 5e-5
    assert _rel(got["3d_vertices"], want["3d_vertices"]) < 5e-5
    assert (got["projected_vertices"].double() - want["projected_vertices"]).abs().max() < 0.25   # input-image pixels (~1e3)
    assert np.abs(got["points"] - want["points"]).max() <= 1              # int truncation of pixel coordinates


def test_predict_batch_matches_oracle(predictor, oracle, cuda_device):
    x = torch.randn(4, 3, 256, 256, generator=torch.Generator().manual_seed(5))
    got = predictor.predict_batch(x, landmark_subset="445")
    want = oracle.predict_batch(x)
    for k in ("3dmm_params", "points", "3d_vertices", "projected_vertices"):
        assert got[k].is_cuda
        assert _rel(got[k], want[k]) < 5e-5, k
    idx = torch.from_numpy(np.load(os.path.join(os.path.dirname(GOLDEN), "..", "dad_3dheads_b200", "assets",
                                                "flame_static.npz"))["keypoints_445"].astype(np.int64))
    assert got["landmarks_445"].shape == (4, 445, 2)
    assert torch.equal(got["landmarks_445"].cpu(), got["projected_vertices"].cpu()[:, idx])


def test_vertex_l2_error_target(predictor, oracle):
    """north_star: vertex L2 error < 1e-4 vs reference (metres, per vertex) through the whole pipeline."""
    x = torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(6))
    got = predictor.predict_batch(x, landmark_subset=None)
    want = oracle.predict_batch(x)
    l2 = (got["3d_vertices"].double().cpu() - want["3d_vertices"]).norm(dim=-1)
    assert l2.max().item() < 1e-4, l2.max().item()


def test_graph_replay_equals_eager(predictor, cuda_device):
    """predict_batch_graphed replays predict_batch from a CUDA graph: bit-identical outputs, for fp32 and raw uint8
    input, and across consecutive calls with different data (static input buffer refreshed every call)."""
    g = torch.Generator().manual_seed(5)
    for make in (lambda: torch.randn(3, 3, 256, 256, generator=g),
                 lambda: torch.randint(0, 256, (3, 256, 256, 3), generator=g, dtype=torch.uint8)):
        for _ in range(2):
            x = make()
            want = {k: v.clone() for k, v in predictor.predict_batch(x).items()}
            got = predictor.predict_batch_graphed(x)
            torch.cuda.synchronize()
            for k in want:
                assert torch.equal(got[k], want[k]), k


def test_config2_batch_512_bf16_encoder(cuda_device):
    """BASELINE configs[2]: batch 512, bf16 encoder, fp32-class FLAME decode + 445-landmark projection.  Size-independent
    checks: finite everywhere, and (eval-mode network: images are independent) a sub-batch run on its own reproduces the
    corresponding rows bit for bit."""
    from dad_3dheads_b200.predictor import FaceMeshPredictor
    pred = FaceMeshPredictor.dad_3dnet(state_dict=synthetic_state_dict(0), precision="bf16")
    x = torch.randn(512, 3, 256, 256, generator=torch.Generator().manual_seed(512))
    out = {k: v.clone() for k, v in pred.predict_batch(x, landmark_subset="445").items()}
    assert out["3d_vertices"].shape == (512, 5023, 3) and out["landmarks_445"].shape == (512, 445, 2)
    assert all(torch.isfinite(v).all() for v in out.values())
    sub = pred.predict_batch(x[300:364], landmark_subset="445")
    for k in out:
        assert torch.equal(out[k][300:364], sub[k]), k
