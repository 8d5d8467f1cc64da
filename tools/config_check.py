"""Run the BASELINE.json configurations at (or near) their stated sizes on one GPU: a few steps
with bit-exact parity on a sample of environments, then timing of one full episode.
    python tools/config_check.py [t1_7x7] [t1_1x1] [syn1024] [syn256]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xroute_env_b200 import VecGame, make_batch, preset_geometry
from oracle.oracle import OracleEnv

CONFIGS = {
    # name: (preset, envs, nets, obs_max_nets, kwargs for the generator, parity sample, steps timed)
    "t1_7x7": ("T1-7x7", 512, 32, -1, {}, 8, 32),            # config 3 per-GPU shard (4096 / 8)
    "t1_1x1": ("T1-1x1", 8192 // 8, 32, -1, {"max_degree": 6}, 8, 32),   # config 5 per-GPU shard
    "syn256": ("SYN-256", 64, 32, -1, {}, 4, 32),            # config 2
    "syn1024": ("SYN-1024", 32, 128, 8, {"hot_spots": 16, "hot_sigma": 32.0}, 2, 8),   # config 4 per-GPU shard
}

def run(name):
    preset, n_envs, n_nets, cap, kw, n_par, n_steps = CONFIGS[name]
    geom = preset_geometry(preset)
    t0 = time.time()
    insts = make_batch(geom, n_envs, n_nets, 777, **kw)
    t_gen = time.time() - t0
    vg = VecGame(geom, insts, device=0, obs_max_nets=cap, pumps_per_sync=int(os.environ.get("XR_PUMPS", "0")))
    vg.reset()
    rng = np.random.default_rng(1)
    orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)   # [n_nets, n_envs]
    oracles = [OracleEnv(geom, insts[e]) for e in range(n_par)]
    # parity on the sampled environments for the first steps
    n_chk = min(6, n_nets)
    for t in range(n_chk):
        vg.step(orders[t])
        delta, done, cum = vg.results_host()
        for e, orc in enumerate(oracles):
            m = orc.step(int(orders[t, e]))
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"], m["overflow"]], (name, t, e)
            oc, oo, ocost = orc.last_paths(); gc, go, gcost = vg.paths(e)
            assert np.array_equal(oc, gc) and np.array_equal(ocost, gcost), (name, t, e)
            if cap < 0:
                assert np.array_equal(vg.obs_host(e).numpy(), orc.obs()), (name, t, e)
    # timing of a fresh episode
    vg.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(n_steps):
        vg.step(orders[t])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {preset} {geom.X}x{geom.Y}x{geom.Z}, {n_envs} envs, {n_nets} nets, obs cap {cap}: parity OK on {n_par} envs x {n_chk} steps; "
          f"{n_steps} steps in {dt*1e3:.1f} ms -> {n_envs*n_steps/dt:.0f} env-steps/s ({dt/n_steps*1e3:.2f} ms/step); "
          f"gen {t_gen:.1f}s; {vg.route_counters()}; {vg.counters()}; mem {torch.cuda.mem_get_info()[0]/2**30:.1f} GiB free", flush=True)
    vg.close()

if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CONFIGS)):
        run(n)
