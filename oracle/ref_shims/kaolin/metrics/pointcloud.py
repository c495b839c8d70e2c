"""``kaolin.metrics.pointcloud.chamfer_distance(p1, p2, w1=1., w2=1., squared=True)`` restated from its documentation:
w1 * mean_i min_j |p1_i - p2_j|^2 + w2 * mean_j min_i |p2_j - p1_i|^2 per batch element (squared distances by default)."""
import torch


def sided_distance(p1, p2):
    d = torch.cdist(p1.double(), p2.double()) ** 2          # [B, N1, N2]
    m, idx = d.min(dim=2)
    return m.to(p1.dtype), idx


def chamfer_distance(p1, p2, w1=1.0, w2=1.0, squared=True):
    d1, _ = sided_distance(p1, p2)
    d2, _ = sided_distance(p2, p1)
    if not squared:
        d1, d2 = d1.sqrt(), d2.sqrt()
    return w1 * d1.mean(dim=-1) + w2 * d2.mean(dim=-1)
