"""Runs the UNMODIFIED reference implementation of the hot path on the CPU (TEST INFRASTRUCTURE ONLY).

What this is: ``activate()`` puts (1) ``oracle/ref_shims`` -- stand-ins for the third-party packages that are neither
installed in the image nor part of /root/reference (smplx, pytorchcv, albumentations, pytorch_toolbelt, hydra, ...; table in
oracle/ref_shims/README.md) -- and (2) the reference's own tree on ``sys.path``: ``/root/reference`` when it exists (the
build container) or its byte-compiled twin ``oracle/_ref`` (built by oracle/build_ref.py; what travels to the GPU box).
After that ``import predictor``, ``from model_training.head_mesh import HeadMesh`` ... load the reference's own files:
  predictor.py:68-211, model_training/head_mesh.py:9-60, model_training/model/flame.py:29-229,
  model_training/model/utils.py:55-101, model_training/model/flame_regression.py:14-106, model_training/model/bifpn.py:11-163,
  model_training/model/encoders.py:9-59, demo.py, demo_utils.py.
The restatements under ``oracle/`` are pinned against these (tests/test_oracle_pinned.py; fixtures from
tools/make_reference_golden.py), and ``bench.py --impl reference`` / the ``cpu_baseline`` leg time them.

Third-party residue of the pin (not reference-owned, restated in the shims): ``smplx.lbs.lbs`` and the ResNet-50 body.

Never imported by the product package.  Do not mix with ``compat/`` in one process (both provide ``predictor`` /
``model_training``): ``activate()`` refuses if those names are already bound to something else.
"""
from __future__ import annotations

import os
import sys
import tempfile
from typing import Any, Dict, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(_HERE, "ref_shims")
REF_SRC = os.environ.get("DAD3D_REFERENCE", "/root/reference")
REF_PYC = os.path.join(_HERE, "_ref")

MODEL_CONFIG = {"backbone": "resnet50", "pretrained": False, "num_filters": 256, "num_channels": 3, "num_classes": 68,
                "img_size": 256, "conv_block": "regular", "limit_value": 3}      # config/model/resnet_regression.yaml
_active_root: Optional[str] = None


def root(prefer: Optional[str] = None) -> Optional[str]:
    """The reference tree to import from: sources if present, else the byte-compiled twin; None if neither exists."""
    order = {"src": [REF_SRC], "pyc": [REF_PYC]}.get(prefer, [REF_SRC, REF_PYC])
    for r in order:
        if os.path.isfile(os.path.join(r, "predictor.py")) or os.path.isfile(os.path.join(r, "predictor.pyc")):
            return r
    return None


def available() -> bool:
    return root() is not None


def kind() -> str:
    r = _active_root or root()
    return "unavailable" if r is None else ("source" if r == REF_SRC else "bytecode")


def activate(prefer: Optional[str] = None) -> str:
    global _active_root
    if _active_root is not None:
        return _active_root
    r = root(prefer)
    if r is None:
        raise RuntimeError("reference harness: neither /root/reference nor oracle/_ref (python oracle/build_ref.py) exists")
    for name in ("predictor", "model_training", "utils", "demo_utils"):
        m = sys.modules.get(name)
        f = getattr(m, "__file__", "") or ""
        if m is not None and not os.path.abspath(f).startswith(os.path.abspath(r)):
            raise RuntimeError(f"reference harness: module {name!r} is already imported from {f}")
    sys.path.insert(0, r)
    sys.path.insert(0, SHIMS)
    _active_root = r
    return r


# ---------------------------------------------------------------------------------------------------------- encoder
def flame_regression(state_dict: Optional[Dict[str, Any]] = None, dtype=None):
    """The reference's ``FlameRegression`` (flame_regression.py:63-106) in eval mode, optionally with ``state_dict`` loaded
    (strict: every key of the released checkpoint's naming must match the module the reference code builds)."""
    activate()
    import torch
    from model_training.model.flame import FLAME_CONSTS
    from model_training.model.flame_regression import FlameRegression
    model = FlameRegression(dict(MODEL_CONFIG), dict(FLAME_CONSTS))
    if state_dict is not None:
        own = model.state_dict()
        sd = dict(state_dict)
        for k in own:                                  # BatchNorm bookkeeping the synthetic initialiser does not carry
            if k.endswith("num_batches_tracked") and k not in sd:
                sd[k] = own[k]
        model.load_state_dict(sd, strict=True)
    if dtype is not None:
        model = model.to(dtype)
    return model.eval()


def trace_checkpoint(state_dict: Dict[str, Any], path: str) -> str:
    """Export like the reference does (train/flame_lightning_model.py:384-401: ``torch.jit.trace`` on a batch of one,
    ``strict=False``) -> a ``.trcd`` the reference predictor ``torch.jit.load``s (predictor.py:72)."""
    import torch
    model = flame_regression(state_dict)
    with torch.no_grad():
        traced = torch.jit.trace(model, torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(0)),
                                 strict=False)
    torch.jit.save(traced, path)
    return path


class cpu_only:
    """Context manager: the reference sees a CPU-only host (``torch.cuda.is_available() -> False``), so its own ``to_device``
    (model/utils.py:31-34) keeps model and tensors on the CPU.  On a GPU box the reference would otherwise run its TorchScript
    module through cuDNN with TF32 convolutions -- a different arithmetic (~1e-3 off its own fp32 CPU path) and not the
    CPU baseline this tier compares against.  The reference code itself stays unmodified."""

    def __enter__(self):
        import torch
        self._orig = torch.cuda.is_available
        torch.cuda.is_available = lambda: False
        return self

    def __exit__(self, *exc):
        import torch
        torch.cuda.is_available = self._orig
        return False


class _CpuPredictor:
    """The reference predictor with every call made under :class:`cpu_only` (attribute access is forwarded)."""

    def __init__(self, inner):
        object.__setattr__(self, "_inner", inner)

    def __call__(self, *a, **k):
        with cpu_only():
            return self._inner(*a, **k)

    def __getattr__(self, name):
        attr = getattr(self._inner, name)
        if callable(attr) and name in ("preprocess", "process", "postprocess"):
            def wrapped(*a, **k):
                with cpu_only():
                    return attr(*a, **k)
            return wrapped
        return attr


def predictor(state_dict: Dict[str, Any], workdir: Optional[str] = None):
    """The reference's ``FaceMeshPredictor`` (predictor.py:68-211) over a checkpoint traced from ``state_dict``, pinned to
    the CPU (see :class:`cpu_only`)."""
    activate()
    import predictor as ref_predictor
    from utils import load_yaml
    config = load_yaml(os.path.join(_active_root, "dad_3dnet.yaml"))
    workdir = workdir or tempfile.mkdtemp(prefix="dad3d_ref_")
    config["model_path"] = trace_checkpoint(state_dict, os.path.join(workdir, "dad_3dheads.trcd"))   # absolute: join keeps it
    with cpu_only():
        return _CpuPredictor(ref_predictor.FaceMeshPredictor(config))


# ---------------------------------------------------------------------------------------------------------- decoder
def head_mesh(flame_config: Optional[Dict[str, int]] = None, image_size: int = 256, dtype=None):
    """The reference's ``HeadMesh`` (head_mesh.py:9-60) over its own ``flame.pkl``; ``dtype=torch.float64`` casts the
    registered buffers for the error yard-stick."""
    activate()
    from model_training.head_mesh import HeadMesh
    hm = HeadMesh(flame_config=flame_config, image_size=image_size)
    if dtype is not None:
        hm = hm.to(dtype)
        hm.flame.dtype = dtype
    return hm


def flame_buffers() -> Dict[str, Any]:
    """What ``FLAMELayer.__init__`` (flame.py:124-180) registers, as numpy arrays -- the ground truth for the packed asset."""
    import numpy as np
    fl = head_mesh().flame
    out = {k: getattr(fl, k).detach().cpu().numpy() for k in ("v_template", "shapedirs", "posedirs", "J_regressor",
                                                              "parents", "lbs_weights", "faces_tensor", "indices_2d")}
    out["faces"] = np.asarray(fl.faces)
    return out
