class Mesh:
    def __init__(self, *a, **k):
        raise NotImplementedError("psbody.mesh is not available (uv_texture demo is outside the hot path)")
