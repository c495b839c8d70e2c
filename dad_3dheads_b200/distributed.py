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
                       group=None, total: Optional[int] = None) -> Dict[str, Tensor]:
    """Concatenate per-rank [B_r, ...] outputs along the batch axis on every rank (rank-major order).

    Ranks may hold different B_r (``shard_range`` gives the first ``n % world`` ranks one extra item): every shard is padded
    to the largest one for the collective and the padding is dropped afterwards; ``total`` = the global item count (default:
    the sum of the per-rank counts, exchanged with one small all-gather)."""
    world = dist.get_world_size(group)
    first = out[next(iter(keys))] if not isinstance(keys, (list, tuple)) else out[keys[0]]
    keys = list(keys)
    dev = first.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([first.shape[0]], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = [int(c) for c in counts.tolist()]
    bmax = max(counts)
    even = all(c == bmax for c in counts)
    res = {}
    for k in keys:
        t = out[k].contiguous()
        if t.shape[0] < bmax:                                  # pad the short shards
            pad = torch.zeros((bmax - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], dim=0)
        buf = buffers.get(k) if buffers else None
        if buf is None or buf.shape[0] != world * bmax:
            buf = torch.empty((world * bmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            if buffers is not None:
                buffers[k] = buf
        dist.all_gather_into_tensor(buf, t, group=group)
        if even:
            res[k] = buf
        else:
            res[k] = torch.cat([buf[r * bmax:r * bmax + counts[r]] for r in range(world)], dim=0)
    return res


class Dad3dComm:
    """The C ABI's own NCCL communicator (include/dad3d.h: dad3d_comm_*): rank 0 creates the NCCL id, torch.distributed (any
    backend) carries its 128 bytes to the other ranks, every rank joins.  ``bcast`` / ``all_gather`` then run inside
    libdad3d.so on the given CUDA stream -- the SURVEY §8(b) exports dad3d_bcast_constants / dad3d_allgather_outputs."""

    def __init__(self, device: torch.device, group=None):
        import ctypes as C

        from . import _lib
        self.lib = _lib.load()
        self._check = _lib.check
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(self.lib.dad3d_comm_unique_id(buf), "dad3d_comm_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        carrier = ident.to(self.device) if dist.get_backend(group) == "nccl" else ident
        dist.broadcast(carrier, 0, group=group)
        raw = bytes(carrier.cpu().tolist())
        h = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(raw)
        _lib.check(self.lib.dad3d_comm_init(C.byref(h), idbuf, self.rank, self.world,
                                            self.device.index if self.device.index is not None else torch.cuda.current_device()),
                   "dad3d_comm_init")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.dad3d_comm_destroy(h)
            except Exception:
                pass
            self._h = None

    def bcast(self, t: Tensor, root: int = 0) -> Tensor:
        assert t.is_cuda and t.is_contiguous()
        self._check(self.lib.dad3d_bcast_constants(self._h, t.data_ptr(), t.numel() * t.element_size(), root,
                                                   torch.cuda.current_stream(t.device).cuda_stream), "dad3d_bcast_constants")
        return t

    def all_gather(self, t: Tensor, out: Optional[Tensor] = None) -> Tensor:
        assert t.is_cuda and t.is_contiguous()
        if out is None:
            out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        assert out.is_contiguous() and out.numel() == self.world * t.numel()
        self._check(self.lib.dad3d_allgather_outputs(self._h, t.data_ptr(), out.data_ptr(), t.numel() * t.element_size(),
                                                     torch.cuda.current_stream(t.device).cuda_stream), "dad3d_allgather_outputs")
        return out
