"""reference path: model_training/head_mesh.py -> dad_3dheads_b200.head_mesh"""
from dad_3dheads_b200.head_mesh import HeadMesh  # noqa: F401
