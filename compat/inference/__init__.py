"""reference path: inference/ -- the PNCC / UV-texture demo helpers (inference/pncc_estimator.py, inference/uv_texture.py)
need the Sim3DR rasteriser and psbody.mesh, which are outside the image -> 3D-head hot path.  When the
reference checkout is on ``sys.path`` BEHIND ``compat/`` its own ``inference`` package is found first only if this directory
is absent; this stub keeps ``import demo_utils`` working without those extras: the two classes import fine and raise a clear
error only when the pncc / uv_texture outputs are actually requested."""
