"""-m gpu, needs >= 2 GPUs (skipped on the single-GPU test box; run with ``gpurun --gpus 2``): world_size-2 NCCL run of the
sharded path.  Checks SURVEY §4(iv) / §8(e): (1) rank r's shard of a batch decoded on its own GPU is BIT-identical to the same
rows of the single-GPU result, (2) the all-gathered buffers hold every rank's shard in rank-major order -- through
torch.distributed and through the C ABI's own NCCL communicator (dad3d_comm_* / dad3d_bcast_constants /
dad3d_allgather_outputs), (3) the pipelined BatchStream with a communication stream returns the same gathered results."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dad_3dheads_b200.distributed import (Dad3dComm, all_gather_outputs, broadcast_flame_static, broadcast_state_dict,
                                                   shard_range)
        from dad_3dheads_b200.encoder_weights import synthetic_state_dict
        from dad_3dheads_b200.flame import load_flame_static
        from dad_3dheads_b200.predictor import DEFAULT_CONFIG, FaceMeshPredictor
        sd = synthetic_state_dict(0 if rank == 0 else 99)            # only rank 0 holds the real constants
        static = load_flame_static()
        sd = broadcast_state_dict(sd, dev)
        static = broadcast_flame_static(static, dev)
        pred = FaceMeshPredictor(dict(DEFAULT_CONFIG), cuda_id=rank, state_dict=sd, precision="fp16x2")
        pred.head_mesh = type(pred.head_mesh)(pred.flame_constants, cuda_id=rank, static=static)
        N = 24
        x = torch.randn(N, 3, 256, 256, generator=torch.Generator().manual_seed(7))
        lo, hi = shard_range(N, rank, world)
        mine = {k: v.clone() for k, v in pred.predict_batch(x[lo:hi], landmark_subset="445").items()}
        keys = ("3dmm_params", "3d_vertices", "landmarks_445")
        gathered = all_gather_outputs(mine, keys)
        ok = True
        if rank == 0:                                                # single-GPU result of the whole batch, same weights
            full = pred.predict_batch(x, landmark_subset="445")
            for k in keys:
                ok = ok and torch.equal(gathered[k], full[k])
        comm = Dad3dComm(dev)
        for k in keys:
            g2 = comm.all_gather(mine[k].contiguous())
            ok = ok and torch.equal(g2, gathered[k])
        t = torch.full((1000,), float(rank + 1), device=dev)
        comm.bcast(t, 0)
        ok = ok and bool((t == 1.0).all())
        # pipelined stream with the gathers on a communication stream (torch.distributed, then the C ABI communicator)
        for cm in (None, comm):
            st = pred.open_stream(x[lo:hi].shape, torch.float32, landmark_subset="445", host_results=False, group=dist.group.WORLD,
                                  comm=cm)
            for _ in range(3):
                st.submit(x[lo:hi].to(dev))
                if st._inflight == st.depth:
                    res = st.collect()
            st.drain()
            res = st.slots[0]["gathered"]
            torch.cuda.synchronize()
            for k in keys:
                ok = ok and torch.equal(res[k], gathered[k])
        q.put((rank, bool(ok)))
    except Exception as e:  # noqa: BLE001 -- report instead of leaving the parent waiting on the queue
        import traceback
        q.put((rank, "error: " + "".join(traceback.format_exception(e))[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_world2_sharded_equals_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] is True for r in res), res
