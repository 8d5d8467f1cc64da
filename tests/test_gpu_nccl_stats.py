"""the statistics all-reduce of the C ABI (xr_stats_allreduce(env, ncclComm_t, stream)) on two
GPUs: every rank steps its own shard of environments, the library sums the int64 statistics vector over a raw NCCL
communicator, and the result equals the sum of the ranks' local vectors.  Skipped on a box with fewer than two GPUs."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _UID(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _worker(rank, world, port, out):
    try:
        _work(rank, world, port, out)
    except Exception as ex:                                     # a failing rank must not leave the other one waiting
        out.put((rank, False, repr(ex), []))
        raise


def _work(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from xroute_env_b200 import VecGame, ispd18_geometry, make_batch
    from xroute_env_b200.dist import shard_range
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nccl = C.CDLL("libnccl.so.2")                              # the copy PyTorch already loaded
    uid = _UID()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.numpy().tobytes()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UID, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    geom = ispd18_geometry(30, 28, 5)
    first, count = shard_range(8, rank, world)
    insts = make_batch(geom, count, 6, seed=500, first_env=first)
    vg = VecGame(geom, insts, device=rank)
    vg.reset()
    rng = np.random.default_rng(rank)
    orders = [list(rng.permutation(i.net_ids)) for i in insts]
    for k in range(6 if rank == 0 else 4):                      # ranks step a different number of times
        vg.step(np.array([o[k] for o in orders], np.int32))
    local = vg.stats().clone().cpu()
    total = vg.stats_allreduce(comm.value).clone().cpu()
    torch.cuda.synchronize()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = bool(torch.equal(total, sum(gathered)))
    out.put((rank, ok, int(total[0]), [int(g[0]) for g in gathered]))
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    vg.close()
    dist.destroy_process_group()


def test_stats_allreduce_over_a_raw_nccl_communicator():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == 4 * 6 + 4 * 4, res                      # env-steps of both shards
