"""Import stub of psbody.mesh (absent; inference/uv_texture.py:5 -- the UV-texture demo is out of scope)."""
