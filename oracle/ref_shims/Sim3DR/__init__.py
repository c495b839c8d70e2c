"""TEST INFRASTRUCTURE.  Stand-in for the reference's Cython package ``Sim3DR`` (Sim3DR/Sim3DR.py:8-29, lib/rasterize.pyx) when
it is not built with Cython: the same two Python functions over the reference's OWN, unmodified C++ rasteriser
(Sim3DR/lib/rasterize_kernel.cpp), compiled by oracle/build_ref.py into oracle/_ref/libsim3dr_ref.so and bound with ctypes
instead of Cython.  Used by the reference's inference/pncc_estimator.py in the CPU arm and as the rasteriser oracle."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "_ref", "libsim3dr_ref.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            raise NotImplementedError("oracle/_ref/libsim3dr_ref.so is not built (python -m oracle.build_ref)")
        _lib = C.CDLL(_SO)
        _lib.sim3dr_ref_rasterize.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_float, C.c_int]
        _lib.sim3dr_ref_get_normal.argtypes = [C.c_void_p] * 3 + [C.c_int] * 2
    return _lib


def _typed(a, dtype, what):
    # the Cython signatures take C-contiguous float32 / int32 buffers and raise on anything else (rasterize.pyx:63-92)
    if a.dtype != dtype or not a.flags.c_contiguous:
        raise ValueError("Buffer dtype mismatch or not C-contiguous: %s" % what)
    return a


def get_normal(vertices, triangles):
    normal = np.zeros_like(vertices, dtype=np.float32)
    _load().sim3dr_ref_get_normal(normal.ctypes.data, _typed(vertices, np.float32, "vertices").ctypes.data,
                                  _typed(triangles, np.int32, "triangles").ctypes.data, vertices.shape[0], triangles.shape[0])
    return normal


def rasterize(vertices, triangles, colors, bg=None, height=None, width=None, channel=None, reverse=False):
    if bg is not None:
        height, width, channel = bg.shape
    else:
        assert height is not None and width is not None and channel is not None
        bg = np.zeros((height, width, channel), dtype=np.uint8)
    buffer = np.zeros((height, width), dtype=np.float32) - 1e8
    if colors.dtype != np.float32:
        colors = colors.astype(np.float32)
    _load().sim3dr_ref_rasterize(_typed(bg, np.uint8, "image").ctypes.data, _typed(vertices, np.float32, "vertices").ctypes.data,
                                 _typed(triangles, np.int32, "triangles").ctypes.data, _typed(colors, np.float32, "colors").ctypes.data,
                                 buffer.ctypes.data, triangles.shape[0], height, width, channel, 1.0, 1 if reverse else 0)
    return bg
