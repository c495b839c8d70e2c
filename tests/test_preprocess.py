"""Pre-processing.  CPU: the numpy restatement of cv2's 8-bit bilinear resize is bit-exact with
the real cv2 (the library the reference calls).  GPU: dad3d_preprocess is bit-exact with the cv2-based host pipeline."""
import numpy as np
import pytest
import torch

SIZES = [(954, 766), (480, 640), (300, 256), (256, 300), (1000, 1000), (257, 255), (123, 77), (64, 48), (256, 256),
         (256, 100), (31, 256), (1, 5)]


def _img(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("hw", SIZES)
def test_resize_oracle_is_bit_exact_with_cv2(hw):
    import cv2
    from oracle.resize_oracle import resize_linear_u8
    h, w = hw
    img = _img(h, w, h * 1000 + w)
    scale = 256 / float(max(h, w))
    nh, nw = max(1, int(round(h * scale))), max(1, int(round(w * scale)))
    if scale == 1.0:
        pytest.skip("no resize")
    ref = cv2.resize(img, dsize=(nw, nh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(resize_linear_u8(img, nh, nw), ref)


def test_letterbox_matches_oracle_transform():
    from dad_3dheads_b200.predictor import letterbox_normalise
    from oracle.predictor_oracle import transform
    for h, w in SIZES[:8]:
        img = _img(h, w, 7)
        assert np.array_equal(letterbox_normalise(img, 256), transform(img, 256))


@pytest.mark.gpu
def test_device_preprocess_bit_exact(cuda_device):
    from dad_3dheads_b200.encoder_weights import synthetic_state_dict
    from dad_3dheads_b200.predictor import FaceMeshPredictor, letterbox_normalise
    pred = FaceMeshPredictor.dad_3dnet(state_dict=synthetic_state_dict(0), precision="bf16")
    sizes = [s for s in SIZES if min(s) >= 2]
    imgs = [_img(h, w, i) for i, (h, w) in enumerate(sizes)]
    got = pred.preprocess_batch(imgs).cpu().numpy()
    for i, im in enumerate(imgs):
        want = np.transpose(letterbox_normalise(im, 256), (2, 0, 1))
        assert np.array_equal(got[i], want), sizes[i]


@pytest.mark.gpu
def test_device_preprocess_same_size_batch(cuda_device):
    """[B,H,W,3] uint8 tensor -> one dad3d_preprocess_batch launch; predict_batch accepts the raw batch directly."""
    import torch
    from dad_3dheads_b200.encoder_weights import synthetic_state_dict
    from dad_3dheads_b200.predictor import FaceMeshPredictor, letterbox_normalise
    pred = FaceMeshPredictor.dad_3dnet(state_dict=synthetic_state_dict(0), precision="fp16x2")
    for h, w in [(256, 256), (300, 411), (97, 64)]:
        imgs = np.stack([_img(h, w, 50 + i) for i in range(3)])
        got = pred.preprocess_batch(torch.from_numpy(imgs).pin_memory()).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i], np.transpose(letterbox_normalise(imgs[i], 256), (2, 0, 1))), (h, w, i)
    raw = torch.from_numpy(np.stack([_img(256, 256, 80 + i) for i in range(2)]))
    a = pred.predict_batch(raw)
    b = pred.predict_batch(pred.preprocess_batch(raw))
    assert torch.equal(a["3dmm_params"], b["3dmm_params"]) and torch.equal(a["3d_vertices"], b["3d_vertices"])
