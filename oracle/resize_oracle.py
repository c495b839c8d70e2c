"""numpy restatement of OpenCV's 8-bit INTER_LINEAR resize (cv::resize fixed-point path; TEST INFRASTRUCTURE ONLY).

Pinned: checked bit-exactly against the real ``cv2.resize`` (installed in this image) in tests/test_preprocess.py --
this is the one stage of the path whose oracle is the actual library the reference calls (through albumentations
LongestMaxSize, predictor.py:198)."""
import numpy as np


def _coeffs(n_dst: int, n_src: int, scale: float, vertical: bool):
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:                          # horizontal taps: out-of-range source columns collapse, fraction zeroed
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= n_src - 1
        f[hi] = 0
        s[hi] = n_src - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64).clip(-32768, 32767)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64).clip(-32768, 32767)
    return np.clip(s, 0, n_src - 1), np.clip(s + 1, 0, n_src - 1), a0, a1   # vertical taps: rows clamped, fraction kept


def resize_linear_u8(src: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    H, W = src.shape[:2]
    scale_x = 1.0 / (np.float64(new_w) / W)
    scale_y = 1.0 / (np.float64(new_h) / H)
    sx, sx1, ax0, ax1 = _coeffs(new_w, W, scale_x, False)
    sy, sy1, ay0, ay1 = _coeffs(new_h, H, scale_y, True)
    S = src.astype(np.int64)
    rows = S[:, sx, :] * ax0[None, :, None] + S[:, sx1, :] * ax1[None, :, None]
    h0, h1 = rows[sy], rows[sy1]
    out = ((((ay0[:, None, None] * (h0 >> 4)) >> 16) + ((ay1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2)
    return out.clip(0, 255).astype(np.uint8)
