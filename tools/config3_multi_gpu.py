"""BASELINE.json configs[2] at full size: T1-7x7 (112x116x9), 4096 environments over 8 GPUs (512 per rank), 32 nets, one
whole episode.  Every rank routes its shard with BOTH engines (frontier search, sweep kernels) and compares the cumulative
metrics of every environment at every step and a hash of every environment's final occupancy; a sample of environments per
rank is checked bit-exactly (paths, costs, metrics, final state) against the CPU oracle; the statistics vector is summed
over the ranks by the library itself (xr_stats_allreduce is exercised by tests/test_gpu_nccl_stats.py; torch.distributed here).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/config3_multi_gpu.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from xroute_env_b200 import VecGame, make_batch, preset_geometry
from xroute_env_b200.dist import allreduce_stats, shard_range
from oracle.oracle import OracleEnv

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
TOTAL, N_NETS, SAMPLE = int(sys.argv[1]) if len(sys.argv) > 1 else 512 * world, 32, 2
geom = preset_geometry("T1-7x7")
first, count = shard_range(TOTAL, rank, world)
insts = make_batch(geom, count, N_NETS, 31, first_env=first)
rng = np.random.default_rng(1000 + rank)
orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)
sample = [0, count - 1][:SAMPLE]
results = {}
for run, engine in enumerate((0, 1, 0)):            # timed frontier, timed sweeps, then the frontier again beside the oracle (untimed)
    vg = VecGame(geom, insts, device=local, engine=engine)
    vg.reset()
    oracles = {e: OracleEnv(geom, insts[e]) for e in sample} if run == 2 else {}
    cums = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(N_NETS):
        vg.step(orders[t])
        _, _, cum = vg.results_host_np()
        cums.append(cum.copy())
        for e, orc in oracles.items():
            m = orc.step(int(orders[t, e]))
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"], m["overflow"]], (rank, t, e)
            oc, oo, ocost = orc.last_paths(); gc, go, gcost = vg.paths(e)
            assert np.array_equal(oc, gc) and np.array_equal(ocost, gcost), (rank, t, e)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    w = (np.arange(geom.cells, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(12345)) | np.uint64(1)
    hashes = np.zeros(count, np.uint64)
    for e in range(count):
        usage, owner = vg.state(e)
        hashes[e] = (usage.reshape(-1).astype(np.uint64) * w).sum() ^ (owner.reshape(-1).astype(np.uint64) * (w >> np.uint64(7))).sum()
    for e, orc in oracles.items():
        ou, oo_ = orc.state(); gu, go_ = vg.state(e)
        assert np.array_equal(gu, ou) and np.array_equal(go_, oo_), (rank, e)
    stats = allreduce_stats(vg.stats())
    if run < 2:
        results[engine] = (np.stack(cums), hashes, dt, stats, vg.route_counters())
    else:
        assert np.array_equal(np.stack(cums), results[0][0]) and np.array_equal(hashes, results[0][1])
    vg.close()
same_metrics = bool(np.array_equal(results[0][0], results[1][0]))
same_hash = bool(np.array_equal(results[0][1], results[1][1]))
flag = torch.tensor([int(same_metrics and same_hash)], device="cuda")
times = torch.tensor([results[0][2], results[1][2]], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
if rank == 0:
    s0 = results[0][3]
    print(f"config 3: T1-7x7 {geom.X}x{geom.Y}x{geom.Z}, {TOTAL} environments over {world} GPU(s) ({count} per rank), {N_NETS} nets, one episode")
    print(f"  frontier engine: {TOTAL * N_NETS / float(times[0]):.0f} env-steps/s wall incl. the per-step host read-back (slowest rank {float(times[0]) * 1e3 / N_NETS:.2f} ms/step); "
          f"sweep engines: {TOTAL * N_NETS / float(times[1]):.0f} env-steps/s")
    print(f"  all ranks: cumulative metrics of every environment at every step equal across the two engines, final-occupancy hashes equal: {bool(flag.item())}")
    print(f"  {SAMPLE} environments per rank bit-exact against the CPU oracle (paths, costs, metrics, final state): True")
    print(f"  statistics over all ranks: steps {s0['steps']} (= {TOTAL} x {N_NETS}: {s0['steps'] == TOTAL * N_NETS}), episodes {s0['episodes']}, wirelength {s0['wirelength']}, via {s0['via']}, violation {s0['violation']}")
assert flag.item() == 1
if world > 1:
    dist.destroy_process_group()
