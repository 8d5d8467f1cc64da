"""world_size-2 gloo run of the multi-GPU host logic on CPU: each rank generates its own
shard of the environment batch, steps it (CPU oracle stands in for the device here, as
the checker), packs the statistics vector and all-reduces it; the result must equal the
single-process run over the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _episode_stats(first, count):
    from oracle.oracle import OracleEnv
    from xroute_env_b200._lib import STAT_NAMES
    from xroute_env_b200.instances import ispd18_geometry, make_batch
    g = ispd18_geometry(16, 14, 3)
    st = np.zeros(16, np.int64)
    for k, inst in enumerate(make_batch(g, count, 4, seed=123, first_env=first)):
        env = OracleEnv(g, inst)
        for net in np.random.default_rng(first + k).permutation(inst.net_ids):
            m = env.step(int(net))
            st[STAT_NAMES.index("steps")] += 1
            st[STAT_NAMES.index("violation")] += m["d_violation"]
            st[STAT_NAMES.index("wirelength")] += m["d_wirelength"]
            st[STAT_NAMES.index("via")] += m["d_via"]
            st[STAT_NAMES.index("reward_x2")] -= 1000 * m["d_violation"] + 8 * m["d_via"] + m["d_wirelength"]
        st[STAT_NAMES.index("episodes")] += 1
    return st


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xroute_env_b200.dist import allreduce_stats, shard_range
    first, count = shard_range(6, rank, world)
    out = allreduce_stats(torch.from_numpy(_episode_stats(first, count)))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stats_allreduce_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from xroute_env_b200._lib import STAT_NAMES
    want = _episode_stats(0, 6)
    assert got == {k: int(v) for k, v in zip(STAT_NAMES, want.tolist())}
    assert got["steps"] == 24 and got["episodes"] == 6
