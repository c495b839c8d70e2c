def imread(path, *a, **k):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    return img[..., ::-1] if img is not None and img.ndim == 3 else img
