"""Step time of the bench workload (or another preset) per routing engine -- development tool.
    python tools/quick_bench.py [preset] [n_envs] [n_nets] [episodes] [engines, e.g. 01]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xroute_env_b200 import VecGame, make_batch, preset_geometry

preset = sys.argv[1] if len(sys.argv) > 1 else "SYN-256"
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_nets = int(sys.argv[3]) if len(sys.argv) > 3 else 32
episodes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
engines = [int(c) for c in (sys.argv[5] if len(sys.argv) > 5 else "01")]
kw = dict(hot_spots=16, hot_sigma=32.0) if preset == "SYN-1024" else ({"max_degree": 6} if preset == "T1-1x1" else {})
obs_cap = 8 if preset == "SYN-1024" else -1
geom = preset_geometry(preset)
insts = make_batch(geom, n_envs, n_nets, 20260000, **kw)
rng = np.random.default_rng(1)
orders = np.stack([np.concatenate([rng.permutation(i.net_ids) for _ in range(episodes + 1)]) for i in insts], 1).astype(np.int32)
ref = None
for eng in engines:
    vg = VecGame(geom, insts, device=0, engine=eng, obs_max_nets=obs_cap, **json.loads(os.environ.get("XR_QB_KW", "{}")))   # e.g. XR_QB_KW='{"window_margin": 22}'
    t = 0
    def episode():
        global t
        vg.reset()
        for _ in range(n_nets):
            vg.step(orders[t]); t += 1
    episode()
    torch.cuda.synchronize()
    c0 = vg.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(episodes):
        episode()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ms = e0.elapsed_time(e1)
    c1 = vg.counters()
    cum = vg.cum.cpu().numpy().copy()
    vg.profile(True)
    t = n_nets
    episode()
    prof = vg.profile_get()
    vg.profile(False)
    steps = episodes * n_nets
    print(f"engine {eng}: {preset} x {n_envs} envs x {n_nets} nets: {ms / steps:.4f} ms/step (wall {1e3 * wall / steps:.4f}), "
          f"{steps * n_envs / (ms / 1e3):.0f} env-steps/s, rounds/step {(c1['relax_passes'] - c0['relax_passes']) / steps:.1f}, "
          f"cells relaxed/step {(c1['cells_relaxed'] - c0['cells_relaxed']) / steps:.0f}, launches/step {(c1['kernel_launches'] - c0['kernel_launches']) / steps:.1f}, "
          f"syncs/step {(c1['host_syncs'] - c0['host_syncs']) / steps:.1f}", flush=True)
    print("   profile (1 episode):", {k: (round(v['ms'], 3), v['launches']) for k, v in prof.items() if v['launches']}, vg.route_counters(), flush=True)
    if ref is None:
        ref = cum
    else:
        print("   cumulative metrics equal to the first engine's:", bool(np.array_equal(ref, cum)), flush=True)
    vg.close()
