"""smplx.utils surface used by the reference: Struct, to_tensor, to_np (flame.py:6,131-180; model/utils.py:2,88)."""
import numpy as np
import torch


class Struct(object):
    def __init__(self, **kwargs):
        for key, val in kwargs.items():
            setattr(self, key, val)


def to_tensor(array, dtype=torch.float32):
    if torch.is_tensor(array):
        return array
    return torch.tensor(array, dtype=dtype)


def to_np(array, dtype=np.float32):
    if "scipy.sparse" in str(type(array)):
        array = array.todense()
    return np.array(array, dtype=dtype)
