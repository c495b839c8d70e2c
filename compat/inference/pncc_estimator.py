class PNCCEstimator:
    def __init__(self, *a, **k):
        raise NotImplementedError("pncc output needs the Sim3DR rasteriser (out of scope here)")
