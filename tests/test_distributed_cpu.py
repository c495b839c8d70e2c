"""-m "not gpu": world_size-2 gloo run of the multi-GPU host logic (constant broadcast, shard ranges, output all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dad_3dheads_b200.distributed import (all_gather_outputs, broadcast_flame_static, broadcast_state_dict,
                                           shard_range)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        g = torch.Generator().manual_seed(100 + rank)            # every rank starts from DIFFERENT values
        sd = {"b.weight": torch.randn(4, 3, generator=g), "a.bias": torch.randn(5, generator=g)}
        static = {"shapedirs": np.full((6, 3, 4), rank + 1, np.float32), "posedirs": np.full((36, 18), rank + 2, np.float32),
                  "v_template": np.full((6, 3), rank + 3, np.float32), "J_regressor": np.full((5, 6), rank + 4, np.float32),
                  "lbs_weights": np.full((6, 5), rank + 5, np.float32), "faces": np.arange(6).reshape(2, 3)}
        sd2 = broadcast_state_dict(sd, dev)
        st2 = broadcast_flame_static(static, dev)
        ref = torch.Generator().manual_seed(100)
        want = {"b.weight": torch.randn(4, 3, generator=ref), "a.bias": torch.randn(5, generator=ref)}
        ok = all(torch.equal(sd2[k], want[k]) for k in want)
        ok = ok and float(st2["shapedirs"].mean()) == 1.0 and float(st2["lbs_weights"].mean()) == 5.0
        ok = ok and np.array_equal(st2["faces"], static["faces"])
        lo, hi = shard_range(7, rank, world)
        full = torch.arange(7 * 2, dtype=torch.float32).reshape(7, 2)
        # all_gather_into_tensor needs equal shards: use a 6-row batch for the gather check
        full6 = torch.arange(6 * 2, dtype=torch.float32).reshape(6, 2)
        l6, h6 = shard_range(6, rank, world)
        got = all_gather_outputs({"x": full6[l6:h6]}, ("x",))["x"]
        ok = ok and torch.equal(got, full6)
        # uneven shards (7 rows over 2 ranks: 4 + 3): padded for the collective, trimmed afterwards
        got7 = all_gather_outputs({"x": full[lo:hi], "y": full[lo:hi] * 2}, ("x", "y"))
        ok = ok and torch.equal(got7["x"], full) and torch.equal(got7["y"], full * 2)
        q.put((rank, ok, (lo, hi), full[lo:hi].shape[0]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_broadcast_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == (0, 4) and res[1][2] == (4, 7)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
