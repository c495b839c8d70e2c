class UVTextureCreator:
    def __init__(self, *a, **k):
        raise NotImplementedError("uv_texture output needs psbody.mesh and the reference's inference/ assets (out of scope here)")
