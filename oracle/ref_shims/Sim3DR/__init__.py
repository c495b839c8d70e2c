"""Import stub of the reference's Cython rasteriser package when it is not built (inference/pncc_estimator.py:3)."""


def rasterize(*a, **k):
    raise NotImplementedError("Sim3DR is not built (pncc demo is outside the hot path)")
