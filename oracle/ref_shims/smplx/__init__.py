"""Shim of smplx==0.1.26 (absent from the image and from /root/reference): only ``smplx.lbs`` and ``smplx.utils``."""
from . import lbs, utils  # noqa: F401
