"""Shim of albumentations==1.0.0 (absent): the three transforms and ``Compose`` the reference's predictor builds
(predictor.py:195-203), implemented over the real ``cv2`` exactly as the package does (functional.py of 1.0.0:
``longest_max_size`` -> ``cv2.resize(INTER_LINEAR)`` with ``py3round``-ed sizes; ``pad_with_params`` ->
``cv2.copyMakeBorder``; ``normalize`` -> float32 ``(img - mean*255) * (1/(std*255))``)."""
import cv2
import numpy as np

from .augmentations.geometric import py3round


class _T:
    def __call__(self, **data):
        data["image"] = self.apply(data["image"])
        return data


class LongestMaxSize(_T):
    def __init__(self, max_size=1024, interpolation=cv2.INTER_LINEAR, always_apply=False, p=1):
        self.max_size, self.interpolation = max_size, interpolation

    def apply(self, img):
        height, width = img.shape[:2]
        scale = self.max_size / float(max(width, height))
        if scale != 1.0:
            new_height, new_width = tuple(py3round(dim * scale) for dim in (height, width))
            img = cv2.resize(img, dsize=(new_width, new_height), interpolation=self.interpolation)
        return img


class PadIfNeeded(_T):
    def __init__(self, min_height=1024, min_width=1024, pad_height_divisor=None, pad_width_divisor=None,
                 border_mode=cv2.BORDER_REFLECT_101, value=None, mask_value=None, always_apply=False, p=1.0):
        self.min_height, self.min_width, self.border_mode, self.value = min_height, min_width, border_mode, value

    def apply(self, img):
        rows, cols = img.shape[:2]
        if rows < self.min_height:
            h_pad_top = int((self.min_height - rows) / 2.0)
            h_pad_bottom = self.min_height - rows - h_pad_top
        else:
            h_pad_top = h_pad_bottom = 0
        if cols < self.min_width:
            w_pad_left = int((self.min_width - cols) / 2.0)
            w_pad_right = self.min_width - cols - w_pad_left
        else:
            w_pad_left = w_pad_right = 0
        return cv2.copyMakeBorder(img, h_pad_top, h_pad_bottom, w_pad_left, w_pad_right, self.border_mode,
                                  value=self.value)


class Normalize(_T):
    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel_value=255.0,
                 always_apply=False, p=1.0):
        self.mean, self.std, self.max_pixel_value = mean, std, max_pixel_value

    def apply(self, img):
        mean = np.array(self.mean, dtype=np.float32)
        mean *= self.max_pixel_value
        std = np.array(self.std, dtype=np.float32)
        std *= self.max_pixel_value
        denominator = np.reciprocal(std, dtype=np.float32)
        img = img.astype(np.float32)
        img -= mean
        img *= denominator
        return img


class Compose:
    def __init__(self, transforms, *args, **kwargs):
        self.transforms = transforms

    def __call__(self, **data):
        for t in self.transforms:
            data = t(**data)
        return data


class _Unsupported:
    def __init__(self, *a, **k):
        raise NotImplementedError("albumentations shim: only the predictor's transforms exist")


BasicTransform = DualTransform = ImageOnlyTransform = Resize = KeypointParams = _Unsupported
