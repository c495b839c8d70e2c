"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL on the GPU box, gloo in CPU tests).

The path is embarrassingly parallel over images (eval-mode BatchNorm: no cross-sample op), so the only
exchanges are (1) a start-up broadcast of the constants from rank 0 (encoder weights, FLAME bases) and (2) an
all-gather of per-image outputs (params, vertices, landmarks) when a caller wants the whole batch on every rank.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

FLAME_BCAST_KEYS = ("shapedirs", "posedirs", "v_template", "J_regressor", "lbs_weights")


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n items over world ranks (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_state_dict(sd: Dict[str, Tensor], device: torch.device, src: int = 0, group=None) -> Dict[str, Tensor]:
    """One packed broadcast of every fp32 tensor of ``sd`` (all ranks must hold tensors of the right shapes; only the
    values of ``src`` survive).  Returns CPU tensors."""
    keys = sorted(sd)
    flat = torch.cat([sd[k].detach().reshape(-1).float() for k in keys]).to(device)
    if dist.get_rank(group) != src:
        flat.zero_()
    dist.broadcast(flat, src, group=group)
    flat = flat.cpu()
    out, off = {}, 0
    for k in keys:
        n = sd[k].numel()
        out[k] = flat[off:off + n].reshape(sd[k].shape).clone()
        off += n
    return out


def broadcast_flame_static(static: Dict[str, np.ndarray], device: torch.device, src: int = 0, group=None,
                           keys: Iterable[str] = FLAME_BCAST_KEYS) -> Dict[str, np.ndarray]:
    out = dict(static)
    for k in keys:
        t = torch.from_numpy(np.ascontiguousarray(static[k])).to(device)
        if dist.get_rank(group) != src:
            t.zero_()
        dist.broadcast(t, src, group=group)
        out[k] = t.cpu().numpy()
    return out


def all_gather_outputs(out: Dict[str, Tensor], keys: Iterable[str], buffers: Optional[Dict[str, Tensor]] = None,
                       group=None) -> Dict[str, Tensor]:
    """Concatenate per-rank [B, ...] outputs along the batch axis on every rank (rank-major order)."""
    world = dist.get_world_size(group)
    res = {}
    for k in keys:
        t = out[k].contiguous()
        buf = buffers.get(k) if buffers else None
        if buf is None:
            buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            if buffers is not None:
                buffers[k] = buf
        dist.all_gather_into_tensor(buf, t, group=group)
        res[k] = buf
    return res
