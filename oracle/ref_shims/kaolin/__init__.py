"""Shim of kaolin (absent): only ``kaolin.metrics.pointcloud.chamfer_distance`` (dad_3dheads_benchmark/utils.py:128,139)."""
