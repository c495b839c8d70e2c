"""-m gpu: the CUDA path against the REFERENCE ITSELF (not the restatement).

On the GPU box /root/reference does not exist; its byte-compiled twin ``oracle/_ref`` (oracle/build_ref.py) does, and
oracle/ref_harness.py runs it on the box's CPU.  Every comparison here is product (libdad3d.so through the C ABI) vs the
unmodified reference code; tolerances are north_star's 1e-4 relative (measured values in the asserts' comments).
The committed-fixture variants of the same checks live in test_flame_gpu.py / test_encoder_gpu.py and run even when
``oracle/_ref`` is missing.
"""
import os
import warnings

import numpy as np
import pytest
import torch

from dad_3dheads_b200.encoder_weights import synthetic_state_dict
from oracle import ref_harness as R
from oracle.flame_oracle import sample_params

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
warnings.filterwarnings("ignore", message="Using torch.cross")


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _contract(got, ref):
    """north_star: element-wise within 1e-4 relative fp32 (absolute floor 1e-4 for values near zero) -> max violation."""
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return ((got - ref).abs() / (1e-4 * ref.abs() + 1e-4)).max().item()


@pytest.fixture(scope="module")
def product(cuda_device):
    from dad_3dheads_b200.predictor import FaceMeshPredictor
    return FaceMeshPredictor.dad_3dnet(state_dict=synthetic_state_dict(0))


@needs_ref
@pytest.mark.parametrize("B", [1, 2, 64, 129])
def test_decode_vs_reference_headmesh(product, cuda_device, B):
    ref = R.head_mesh()
    p = sample_params(B, seed=40 + B)
    want_v = ref.vertices_3d(p.clone())
    q = p.clone()
    want_p = ref.reprojected_vertices(q, to_2d=True)
    for hilo, tol, l2 in ((False, 5e-5, 1e-4), (True, 2e-6, 1e-6)):      # default one-product kernel / strict hi-lo blend
        v3, pj = product.head_mesh.decode(p.to(cuda_device), to_2d=True, hilo=hilo)
        assert _rel(v3, want_v) < tol and _rel(pj, want_p) < tol, (hilo, _rel(v3, want_v), _rel(pj, want_p))
        assert _contract(v3, want_v) < 1.0 and _contract(pj, want_p) < 1.0
        assert (v3.cpu() - want_v).norm(dim=-1).max().item() < l2        # vertex L2 (m); north_star target < 1e-4
    # the reference-facing methods on CPU tensors, side effect included
    q2 = p.clone()
    got_p = product.head_mesh.reprojected_vertices(q2, to_2d=True)
    assert torch.equal(q2, q) and _rel(got_p, want_p) < 2e-6


@needs_ref
def test_encoder_vs_reference_flame_regression(cuda_device):
    from dad_3dheads_b200.encoder import Dad3dEncoder
    sd = synthetic_state_dict(4)
    m = R.flame_regression(sd, dtype=torch.float64)
    x = torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(31))
    with torch.no_grad():
        want = m(x.double())
    for mode, tol in (("fp32", 3e-5), ("fp16x2", 3e-5)):
        got = Dad3dEncoder(sd, cuda_device, precision=mode)(x.to(cuda_device))
        for k in want:
            assert _rel(got[k], want[k]) < tol, (mode, k)
        assert _contract(got["OUTPUT_3DMM_PARAMS"], want["OUTPUT_3DMM_PARAMS"]) < 1.0, mode


@needs_ref
def test_predictor_call_vs_reference_predictor(product):
    """FaceMeshPredictor.__call__ (predictor.py:78-83) on the demo image and on odd sizes: same keys, dtypes, shapes,
    in-place semantics; values within the contract; integer landmark pixels within 1."""
    import cv2
    ref = R.predictor(synthetic_state_dict(0))
    imgs = [cv2.cvtColor(cv2.imread(os.path.join(GOLDEN, "demo_head_1.jpeg")), cv2.COLOR_BGR2RGB)]
    g = np.random.default_rng(1)
    imgs += [g.integers(0, 256, s + (3,), dtype=np.uint8) for s in ((300, 517), (641, 203), (256, 256))]
    for img in imgs:
        want, got = ref(img.copy()), product(img.copy())
        assert set(got) == set(want)
        for k in want:
            assert tuple(got[k].shape) == tuple(want[k].shape), (k, got[k].shape, want[k].shape)
            if torch.is_tensor(want[k]):
                assert torch.is_tensor(got[k]) and got[k].dtype == want[k].dtype and got[k].device == want[k].device, k
            else:
                assert isinstance(got[k], np.ndarray) and got[k].dtype.kind == want[k].dtype.kind, (k, got[k].dtype)
        errs = {k: _rel(got[k], want[k]) for k in ("3dmm_params", "3d_vertices", "projected_vertices")}
        dpx = int(np.abs(got["points"] - want["points"]).max())
        assert all(v < 5e-5 for v in errs.values()) and dpx <= 1, (img.shape, errs, dpx)


def test_predictor_fixture(product):
    """Same check against the committed output of the reference predictor (tests/golden/reference_predictor.npz)."""
    import cv2
    z = np.load(os.path.join(GOLDEN, "reference_predictor.npz"))
    img = cv2.cvtColor(cv2.imread(os.path.join(GOLDEN, "demo_head_1.jpeg")), cv2.COLOR_BGR2RGB)
    got = product(img)
    assert _rel(got["3dmm_params"], z["params_3dmm"]) < 5e-5
    assert _rel(got["3d_vertices"], z["vertices_3d"]) < 5e-5
    assert _rel(got["projected_vertices"], z["projected_vertices"]) < 5e-5
    assert np.abs(got["points"] - z["points"]).max() <= 1
