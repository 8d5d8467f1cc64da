"""Workload for ncu captures: bench configuration, 16 steps into an episode (16 nets left in
every environment), then isolated launches of the observation and metrics kernels.
The LAST k_obs / k_metrics launches in the capture are the isolated ones:
algorithmic bytes = 64 envs * 4 B * (2 + 7*16) * 589824 cells = 17.21 GB (obs),
64 * 4 B * 589824 = 0.151 GB (metrics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xroute_env_b200 import VecGame, make_batch, preset_geometry
from bench import make_orders, N_NETS, SEED, ENVS_PER_GPU

geom = preset_geometry("SYN-256")
insts = make_batch(geom, ENVS_PER_GPU, N_NETS, SEED)
sched = make_orders(insts, 40, SEED)
vg = VecGame(geom, insts, device=0, metrics_mode=1)
vg.reset()
for t in range(16):
    vg.step(sched[t])
print("obs", vg.kernel_bench("obs", 2))
print("metrics", vg.kernel_bench("metrics", 2))
for t in range(16, 20):
    vg.step(sched[t])
torch.cuda.synchronize()
print(vg.route_counters())
vg.close()
