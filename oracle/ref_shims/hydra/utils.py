import importlib
import os


def get_original_cwd():
    return os.getcwd()


def instantiate(config, *args, **kwargs):
    """Minimal ``_target_`` instantiation (what model_training/model/__init__.py:7 needs)."""
    cfg = dict(config)
    target = cfg.pop("_target_")
    module, name = target.rsplit(".", 1)
    cls = getattr(importlib.import_module(module), name)
    cfg.update(kwargs)
    return cls(*args, **cfg)
