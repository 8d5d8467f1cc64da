"""Diagnostics: route-path usage and per-step time of the bench workload for a few settings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xroute_env_b200 import VecGame, make_batch, preset_geometry
from bench import make_orders, N_NETS, SEED

geom = preset_geometry("SYN-256")
N = int(os.environ.get("N_ENVS", 64))
insts = make_batch(geom, N, N_NETS, SEED)
sched = make_orders(insts, 64, SEED)
settings = [eval(a) for a in sys.argv[1:]] or [dict()]
for kw in settings:
    vg = VecGame(geom, insts, device=0, **kw)
    ts = []
    for t in range(64):
        if t % N_NETS == 0:
            vg.reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        vg.step(sched[t])
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    vg.profile(True)
    for t in range(32):
        if t % N_NETS == 0:
            vg.reset()
        vg.step(sched[t])
    tl = vg.debug_timeline()
    prof = vg.profile_get()
    print(kw, "ms/step mean %.2f median %.2f max %.2f" % (1e3*np.mean(ts), 1e3*np.median(ts), 1e3*np.max(ts)),
          vg.route_counters(), vg.counters(), vg.debug_counters())
    print("   ", {k: (round(v["ms"],1), v["launches"]) for k, v in prof.items() if v["launches"]})
    print("    timeline(ms from step start):", tl)
    vg.close()
