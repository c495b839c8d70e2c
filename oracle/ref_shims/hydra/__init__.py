"""Import stub of hydra (absent): model_training/utils.py:4 and model_training/model/__init__.py:1 import it at module
level; nothing on the image -> 3D-head path calls it except ``instantiate`` for ``_target_`` configs."""
from . import utils  # noqa: F401


def main(*a, **k):
    def deco(fn):
        return fn
    return deco
