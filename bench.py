#!/usr/bin/env python
"""bench.py -- env-steps/s of the XRoute hot path (obs build + maze route + reward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload of the headline line (BASELINE.json configs[1]): synthetic 256x256x9 grid, 64 environments per GPU,
32 nets per environment, random net ordering, episodes back to back (reset every 32 steps).  A "step" is one
batched environment step: every environment of the batch routes one net, updates its metrics and its
observation.  Weak scaling: each rank owns its own 64-environment shard (seeded by global environment id), no
data-path collective; the episode statistics vector is all-reduced once after the timed region.

The same JSON line carries, at N = 1, the other configurations of BASELINE.json as bounded legs
(`legs`: T1-7x7 x 512 envs = one GPU's shard of configs[2] next to the CPU path on ALL host cores on the same grid,
T1-1x1 x 1024 envs = configs[4], SYN-1024 x 32 envs = configs[3]; `ispd18_test1`: a real 7x7-gcell region), a longer
self-timed run (`long_run`, 10 episodes) and the roofline block (dominant kernel + the HBM-bound kernels the north
star names).  Prints ONE JSON line on rank 0.  `--impl reference` times the CPU oracle port of the same path on all
host cores (the reference's own router is an absent external binary).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESET = "SYN-256"
ENVS_PER_GPU = 64
N_NETS = 32
SEED = 20260000
GEN_KW = {"SYN-1024": {"hot_spots": 16, "hot_sigma": 32.0}, "T1-1x1": {"max_degree": 6}}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _kernel_facts():
    """ncu-derived facts of the final build (profiles/r2_kernel_facts.json, written by profiles/summarize.py facts
    from the committed captures): DRAM traffic per launch, issue-slot utilisation, SM coverage."""
    p = os.path.join(ROOT, "profiles", "r2_kernel_facts.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/xr_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], 0, set()
        if os.path.getsize(self.path) == 0:          # the sampler never got a line out: one synchronous query
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=20).stdout
                with open(self.path, "w") as f:
                    f.write(out)
            except Exception:
                pass
        with open(self.path) as f:
            for line in f:
                c = [s.strip() for s in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx = max(mx, float(c[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_orders(insts, n_steps, seed, n_nets=N_NETS):
    """Per-environment action schedule: a fresh random permutation of the net ids per episode."""
    rng = np.random.default_rng(seed)
    n_eps = (n_steps + n_nets - 1) // n_nets + 1
    sched = np.zeros((n_steps + n_nets, len(insts)), np.int32)
    for e, inst in enumerate(insts):
        ids = np.array(inst.net_ids, np.int32)
        assert len(ids) == n_nets
        seq = np.concatenate([rng.permutation(ids) for _ in range(n_eps)])
        sched[:, e] = seq[: sched.shape[0]]
    return sched


# --------------------------------------------------------------------------- CPU side
def _cpu_worker(args):
    """One process = one environment at a time (how the reference runs), oracle port: observation build + maze
    route + metrics + reward per env-step."""
    preset, n_nets, first_env, n_envs, budget_s = args
    import ctypes as C
    from oracle.oracle import OracleEnv, lib
    from xroute_env_b200.instances import make_instance, preset_geometry
    geom = preset_geometry(preset)
    buf = np.empty((2 + 7 * n_nets, geom.cells), np.float32)
    steps = settled = 0
    t0 = time.perf_counter()
    for e in range(first_env, first_env + n_envs):
        inst = make_instance(geom, n_nets, SEED + e, **GEN_KW.get(preset, {}))
        env = OracleEnv(geom, inst)
        order = np.random.default_rng(SEED + e).permutation(inst.net_ids)
        lib().orc_obs(env._h, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.shape[0])
        for net in order:
            env.step(int(net))
            lib().orc_obs(env._h, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.shape[0])
            steps += 1; settled += env.settled()
            if time.perf_counter() - t0 > budget_s:
                return steps, settled, time.perf_counter() - t0
    return steps, settled, time.perf_counter() - t0


def _cpu_region_worker(args):
    """The same for the real ispd18_test1 region (every environment is the same region, another net order)."""
    name, first_env, n_envs, budget_s = args
    import ctypes as C
    from oracle.oracle import OracleEnv, lib
    from xroute_env_b200.ispd import load_regions
    geom, inst = load_regions(os.path.join(ROOT, "tests", "golden", "ispd18_test1_regions.npz"))[name]
    nets = inst.net_ids
    buf = np.empty((2 + 7 * len(nets), geom.cells), np.float32)
    env = OracleEnv(geom, inst)
    steps = 0
    t0 = time.perf_counter()
    for e in range(first_env, first_env + n_envs):
        env.reset()
        for net in np.random.default_rng(SEED + e).permutation(nets):
            env.step(int(net))
            lib().orc_obs(env._h, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.shape[0])
            steps += 1
            if time.perf_counter() - t0 > budget_s:
                return steps, 0, time.perf_counter() - t0
    return steps, 0, time.perf_counter() - t0


def _pool_cores(bytes_per_worker: float) -> int:
    cores = os.cpu_count() or 1
    try:
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available / max(bytes_per_worker, 1.0))))
    except Exception:
        pass
    return cores


def cpu_run(cores: int, budget_s: float, preset: str = PRESET, n_nets: int = N_NETS, region: str | None = None):
    """Oracle port on `cores` processes for about `budget_s` seconds.  Returns (env-steps/s, steps, cells settled/s, wall)."""
    fn = _cpu_region_worker if region else _cpu_worker
    mk = (lambda i: (region, 64 * i, 64, budget_s)) if region else (lambda i: (preset, n_nets, 64 * i, 64, budget_s))
    if cores == 1:
        res = [fn(mk(0))]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(fn, [mk(i) for i in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return steps / wall, steps, sum(r[1] for r in res) / wall, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    cores = _pool_cores(1.5e9)                 # each worker holds a 0.53 GB observation buffer
    per_step = []
    total_steps = 0
    # each "step" of this arm is a bounded sample: every core works for ~budget seconds
    budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        v, steps, settled_s, wall = cpu_run(cores, budget)
        if i >= args.warmup:
            per_step.append((v, steps, wall, settled_s))
            total_steps += steps
    v = sum(p[1] for p in per_step) / sum(p[2] for p in per_step)
    geom_dims = "256x256x9"
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": v, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.mean(p[2] for p in per_step), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"SYN-256 {geom_dims}, {N_NETS} nets/env, random net order, obs+route+reward per env-step",
                   "grid": geom_dims, "nets_per_env": N_NETS,
                   "note": "CPU oracle port (oracle/xr_oracle.c), one environment per process on every host core; "
                           "the reference's own router is an external binary that is not in its tree"},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{total_steps} env-steps in {len(per_step)} samples of ~{budget:.0f}s on {cores} processes"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cells_settled_per_s": statistics.mean(p[3] for p in per_step),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU legs
def _episodes(vg, orders, n_nets, episodes, read_results=False):
    """Time `episodes` back-to-back episodes (reset + n_nets steps) with CUDA events after one warm-up episode."""
    import torch
    t = 0

    def episode():
        nonlocal t
        vg.reset()
        for _ in range(n_nets):
            vg.step(orders[t]); t += 1
            if read_results:
                vg.results_host_np()
    episode()
    torch.cuda.synchronize()
    c0 = vg.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(episodes):
        episode()
    e1.record()
    torch.cuda.synchronize()
    c1 = vg.counters()
    return e0.elapsed_time(e1), {k: c1[k] - c0[k] for k in c0}


def synthetic_leg(device, preset, n_envs, n_nets, episodes, obs_cap=-1, cpu_budget=0.0, steps=None):
    """One BASELINE.json configuration at its single-GPU size: env-steps/s of whole episodes (device-timed, results
    read back to the host every step), optionally next to the CPU path on one and on all host cores."""
    import torch
    from xroute_env_b200 import VecGame, make_batch, preset_geometry
    geom = preset_geometry(preset)
    t0 = time.perf_counter()
    insts = make_batch(geom, n_envs, n_nets, SEED, **GEN_KW.get(preset, {}))
    t_gen = time.perf_counter() - t0
    vg = VecGame(geom, insts, device=device, obs_max_nets=obs_cap)
    n_steps = n_nets if steps is None else steps
    orders = make_orders(insts, (episodes + 1) * n_nets, SEED + 5, n_nets)
    ms, dc = _episodes(vg, orders, n_steps, episodes, read_results=True)
    out = {"grid": f"{preset} {geom.X}x{geom.Y}x{geom.Z}", "envs": n_envs, "nets_per_env": n_nets,
           "value": episodes * n_steps * n_envs / (ms / 1e3), "unit": "env-steps/s", "ms_per_step": ms / (episodes * n_steps),
           "timed": f"{episodes} episode(s) of {n_steps} steps after one warm-up episode; reset + step + results read back to "
                    "pinned host memory every step; CUDA events",
           "cells_relaxed_per_s": dc["cells_relaxed"] / (ms / 1e3), "route_paths": vg.route_counters(),
           "instance_generation_s": round(t_gen, 1)}
    if obs_cap >= 0:
        out["obs_max_nets"] = obs_cap
    vg.close()
    del vg
    torch.cuda.empty_cache()
    if cpu_budget > 0:
        from oracle import oracle
        oracle.build()
        v1, s1, _, w1 = cpu_run(1, cpu_budget / 2, preset, n_nets)
        cores = _pool_cores(4.0 * (2 + 7 * n_nets) * geom.cells * 3)
        va, sa, _, wa = cpu_run(cores, cpu_budget, preset, n_nets)
        out["cpu_1_thread"] = {"value": v1, "unit": "env-steps/s", "sample": f"{s1} env-steps in {w1:.1f}s"}
        out["cpu_all_cores"] = {"value": va, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{sa} env-steps in {wa:.1f}s, one environment per process "
                                          "(oracle/xr_oracle.c: obs + route + reward)"}
        out["x_over_cpu_all_cores"] = out["value"] / va
        out["x_over_cpu_1_thread"] = out["value"] / v1
    return out


def ispd_leg(device: int, cpu: bool, n_envs: int = 512, episodes: int = 2):
    import torch
    from xroute_env_b200 import VecGame
    from xroute_env_b200.ispd import load_regions
    name = "t1_7x7_y79800"
    geom, inst = load_regions(os.path.join(ROOT, "tests", "golden", "ispd18_test1_regions.npz"))[name]
    nets = inst.net_ids
    rng = np.random.default_rng(SEED)
    orders = np.stack([np.concatenate([rng.permutation(nets) for _ in range(episodes + 1)]) for _ in range(n_envs)], 1)
    orders = np.ascontiguousarray(orders, np.int32)
    vg = VecGame(geom, [inst] * n_envs, device=device)
    ms, dc = _episodes(vg, orders, len(nets), episodes, read_results=True)
    out = {"region": f"{name} (routeBox 39900,79800-79800,119700; {geom.X}x{geom.Y}x{geom.Z}, {len(nets)} nets, "
                     f"{len(inst.ap_net)} access points, {len(inst.block_xyz)} blockages)",
           "envs": n_envs, "value": episodes * len(nets) * n_envs / (ms / 1e3), "unit": "env-steps/s",
           "ms_per_step": ms / (episodes * len(nets)), "route_paths": vg.route_counters()}
    vg.close()
    # the single-environment legacy API on the same region: Game.reset/step as train_PPO.py calls them, every
    # observation copied to a fresh CPU tensor (what the reference's Game returns)
    from xroute_env_b200 import Game
    game = Game(geometry=geom, instances=[inst], device=device, max_nets=max(nets), max_aps=len(inst.ap_net))
    game.reset()
    for net in orders[:4, 0]:
        game.step(int(net))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obs0, _ = game.reset()
    d2h = obs0.numel() * 4
    for net in orders[: len(nets), 1]:
        o, done, *_ = game.step(int(net))
        d2h += o.numel() * 4
    wall = time.perf_counter() - t0
    game._vec.close()
    out["legacy_game_api"] = {"value": len(nets) / wall, "unit": "env-steps/s", "ms_per_step": 1e3 * wall / len(nets),
                              "d2h_bytes_per_step": d2h // (len(nets) + 1), "d2h_gbs": d2h / wall / 1e9,
                              "note": "xroute_env_b200.Game (reference-compatible single environment): one episode, "
                                      "observation tensors copied to the host every step"}
    if cpu:
        from oracle import oracle
        oracle.build()
        v1, s1, _, w1 = cpu_run(1, 5.0, region=name)
        cores = _pool_cores(4.0 * (2 + 7 * len(nets)) * geom.cells * 3)
        va, sa, _, wa = cpu_run(cores, 8.0, region=name)
        out["cpu_1_thread"] = {"value": v1, "unit": "env-steps/s", "sample": f"{s1} env-steps in {w1:.1f}s"}
        out["cpu_all_cores"] = {"value": va, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{sa} env-steps in {wa:.1f}s, one environment per process"}
        out["x_over_cpu_all_cores"] = out["value"] / va
        out["x_over_cpu_1_thread"] = out["value"] / v1
    return out


def _obs_update_bytes(insts, sched_rows):
    """Algorithmic bytes of the in-place observation update (k_obs_update) for the given steps: every access point of a net
    that shifts down one block is cleared at the old block and set at the new one (1 + 6 * adjacent floats each), the routed
    net's block is cleared, the order channel rewritten from the routed net's rank on."""
    total = 0.0
    per_env = []
    for inst in insts:
        cells = set(map(tuple, inst.ap_xyz.tolist()))
        netof = {tuple(c): int(n) for c, n in zip(inst.ap_xyz.tolist(), inst.ap_net)}
        w = {}
        for c, n in netof.items():
            adj = any(netof.get((c[0] + dx, c[1] + dy, c[2] + dz)) == n
                      for dx, dy, dz in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)))
            w[n] = w.get(n, 0) + (7 if adj else 1)
        per_env.append(w)
    remaining = [set(w) for w in per_env]
    for row in sched_rows:
        for e, a in enumerate(row):
            a = int(a)
            if a not in remaining[e]:              # (a new episode started)
                remaining[e] = set(per_env[e])
            later = [n for n in remaining[e] if n > a]
            total += 4.0 * (per_env[e][a] + 2 * sum(per_env[e][n] for n in later) + len(later) + 1)
            remaining[e].discard(a)
    return total


# --------------------------------------------------------------------------- GPU side
def run_ours(args):
    import torch
    import torch.distributed as dist
    from xroute_env_b200 import VecGame, make_batch, preset_geometry
    from xroute_env_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's start-up banner (printed by the library on its first
        # communicator when NCCL_DEBUG is set on the box) goes to stderr with everything else
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    geom = preset_geometry(PRESET)
    K, W = args.steps, args.warmup
    LONG_EPISODES = 10
    insts = make_batch(geom, ENVS_PER_GPU, N_NETS, SEED, first_env=rank * ENVS_PER_GPU)
    vg = VecGame(geom, insts, device=local)
    total_steps = W + 3 * K + (8 + LONG_EPISODES) * N_NETS
    sched = make_orders(insts, total_steps, SEED + 17 * rank)
    pinned = torch.from_numpy(sched).pin_memory()
    sched_p = pinned.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"t": 0}

    def one_step(read_results: bool):
        t = state["t"]
        if t % N_NETS == 0:
            vg.reset()
        vg.step(sched_p[t])
        state["t"] = t + 1
        if read_results:
            delta, done, cum = vg.results_host_np()
            return float(-(500.0 * delta[:, 0].sum() + 4.0 * delta[:, 2].sum() + 0.5 * delta[:, 1].sum()))
        return None

    def timed(n, read_results):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            one_step(read_results)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            state["rank_ms"] = [float(v.item()) for v in allt]
            ms = max(state["rank_ms"])                    # max over ranks
        else:
            state["rank_ms"] = [ms]
        return ms

    def align():
        while state["t"] % N_NETS != 0:
            one_step(False)

    # clocks are sampled from before the warm-up to the end of the timed legs (the timed region alone can be
    # shorter than nvidia-smi's start-up when eight ranks share the host)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):                                       # warm-up (untimed)
        one_step(False)
    align()                                                  # every timed region starts at an episode boundary

    c0 = vg.counters()
    ms_dev = timed(K, read_results=False)
    rank_ms = [v / K for v in state["rank_ms"]]
    c1 = vg.counters()
    align()
    ms_e2e = timed(K, read_results=True)
    align()
    # a longer self-timed run through the public API (results read back every step): 10 whole episodes
    ms_long = timed(LONG_EPISODES * N_NETS, read_results=True)
    long_rank_ms = [v / (LONG_EPISODES * N_NETS) for v in state["rank_ms"]]
    clocks = sampler.stop()
    align()
    # profiled leg: per-kernel-class CUDA-event timing on the launching stream
    vg.profile(True)
    t_prof0 = state["t"]
    p0 = vg.counters()
    ms_prof = timed(K, read_results=False)
    prof = vg.profile_get()
    vg.profile(False)
    p1 = vg.counters()

    # isolated (burst) timing of the two HBM-bound kernels the north star names, 16 steps into an episode (16 nets left in
    # every environment: the launch shape of tools/profile_kernels.py, whose ncu capture gives `traffic`)
    align()
    for _ in range(N_NETS // 2):
        one_step(False)
    iso = {k: vg.kernel_bench(k, 10) for k in ("obs", "metrics")}

    env_steps = K * ENVS_PER_GPU * world
    value = env_steps / (ms_dev / 1e3)
    e2e = env_steps / (ms_e2e / 1e3)
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    cells_relaxed = c1["cells_relaxed"] - c0["cells_relaxed"]
    relax_passes = c1["relax_passes"] - c0["relax_passes"]
    if world > 1:
        t = torch.tensor([cells_relaxed, relax_passes], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        cells_relaxed_all, relax_passes_all = float(t[0]), float(t[1])
    else:
        cells_relaxed_all, relax_passes_all = float(cells_relaxed), float(relax_passes)

    route_paths = vg.route_counters()
    # episode statistics: the only collective on this path (sum of a tiny int64 vector)
    stats = vg.stats().clone()
    if world > 1:
        dist.all_reduce(stats)
    stats = {k: int(v) for k, v in zip(_lib.STAT_NAMES, stats.cpu().tolist())}

    # ---- kernels of the profiled leg (rank 0) and the roofline block
    peak, peak_src = _peaks()
    facts = _kernel_facts()
    cells = geom.cells
    kern = {}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    prof_cells = p1["cells_relaxed"] - p0["cells_relaxed"]
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        kern[k] = {"ms_total": round(v["ms"], 3), "launches": int(v["launches"]), "share": round(v["ms"] / tot_ms, 4),
                   "avg_us": round(1e3 * v["ms"] / v["launches"], 2)}
    if "obs" in kern:       # per step the observation kernel is the in-place update (+ the per-episode reset passes)
        upd_bytes = _obs_update_bytes(insts, sched_p[t_prof0:t_prof0 + K])
        kern["obs"]["algorithmic_bytes"] = upd_bytes
        kern["obs"]["achieved_gbs"] = round(upd_bytes / (prof["obs"]["ms"] / 1e3) / 1e9, 2)
        kern["obs"]["note"] = ("k_obs_update: scattered 4-byte stores at the access points of the nets whose block shifts (latency-bound, "
                               "not a stream); the time also holds the reset passes of the episode boundaries in the leg; the streaming "
                               "full build is roofline.hbm_kernels.k_obs")
    dom = max((k for k in kern if k.startswith("route") or k.startswith("sweep")), key=lambda k: kern[k]["ms_total"], default=None)
    roofline = {"peak": peak, "unit": "GB/s", "peak_source": peak_src}
    if dom is not None:
        # section 8(d)'s yardstick for the maze kernels: 12 B per cell relaxed (word read + conditional write + cost), all
        # route kernels of the leg together
        route_ms = sum(kern[k]["ms_total"] for k in kern if k.startswith("route") or k.startswith("sweep"))
        alg = 12.0 * prof_cells
        fr = facts.get("k_route_frontier", {})
        fw = facts.get("k_route_win", {})
        n_route_launch = sum(kern[k]["launches"] for k in kern if k.startswith("route") or k.startswith("sweep"))
        roofline.update({
            "kernel": f"route kernels ({sum(kern[k]['share'] for k in kern if k.startswith('route') or k.startswith('sweep')):.3f} of the kernel time of the "
                      f"profiled leg; largest share: {'k_route_frontier' if dom == 'route_frontier' else 'k_' + dom}).  k_route_frontier is one launch per step and the step's "
                      "critical path; the window kernels (k_route_win<C>, hybrid policy) run beside it on their own streams, so kernel-time shares add up to more "
                      "than the step",
            "shares": {('k_' + k): kern[k]["share"] for k in kern if k.startswith("route") or k.startswith("sweep")},
            "avg_launch_us_by_kernel": {('k_' + k): kern[k]["avg_us"] for k in kern if k.startswith("route") or k.startswith("sweep")},
            "bound": "hbm", "achieved": round(alg / (route_ms / 1e3) / 1e9, 2), "frac": round(alg / (route_ms / 1e3) / 1e9 / peak, 5),
            "algorithmic_bytes_per_launch": alg / max(1, n_route_launch), "avg_launch_us": round(1e3 * route_ms / max(1, n_route_launch), 2),
            "traffic": fr.get("dram_bytes_per_launch"),
            "traffic_source": (fr.get("source") or "") + " -- one k_route_frontier launch; window_kernel holds the same facts of one k_route_win<4> launch",
            "window_kernel": {k: fw.get(k) for k in ("source", "dram_bytes_per_launch", "issue_slots_busy_pct", "sm_busy_pct", "grid", "block", "registers")},
            "efficiency": {"cells_relaxed_per_s": prof_cells / (route_ms / 1e3), "of_5.44e11_cells_per_s": prof_cells / (route_ms / 1e3) / 5.44e11,
                           "issue_slots_busy_pct": fr.get("issue_slots_busy_pct"), "sm_busy_pct": fr.get("sm_busy_pct"),
                           "ctas": fr.get("grid"), "threads_per_cta": fr.get("block"), "registers": fr.get("registers"),
                           "ncu_source": fr.get("source")},
            "note": "the route kernel is a latency-bound sparse search (rounds of one L2 round trip each), not a stream: the goal-directed "
                    "search relaxes ~8x fewer cells than the sweeps it replaces, so the section 8(d) yardstick (12 B x cells relaxed) is a small "
                    "fraction of the HBM peak by construction; the HBM-bound kernels of the path are below",
        })
    hbm = {}
    fo, fm = facts.get("k_obs", {}), facts.get("k_metrics", {})
    hbm["k_obs"] = {"what": "full observation build, isolated, all 64 environments with 16 nets left", "achieved": round(iso["obs"]["gbs"], 1),
                    "frac": round(iso["obs"]["gbs"] / peak, 4), "algorithmic_bytes_per_launch": iso["obs"]["bytes"],
                    "avg_launch_us": round(1e3 * iso["obs"]["ms"], 2), "traffic": fo.get("dram_bytes_per_launch"),
                    "traffic_source": fo.get("source")}
    hbm["k_metrics"] = {"what": "congestion scan of the occupancy field (reward kernel), isolated, all 64 environments",
                        "achieved": round(iso["metrics"]["gbs"], 1), "frac": round(iso["metrics"]["gbs"] / peak, 4),
                        "algorithmic_bytes_per_launch": iso["metrics"]["bytes"], "avg_launch_us": round(1e3 * iso["metrics"]["ms"], 2),
                        "traffic": fm.get("dram_bytes_per_launch"), "traffic_source": fm.get("source")}
    roofline["hbm_kernels"] = hbm
    roofline["isolated_method"] = "CUDA events around 10 back-to-back launches on the launching stream (xr_kernel_bench)"

    # ---- the same workload with (a) a full observation rebuild every step and (b) the metric scan every step: the HBM-bound
    # form of the step, one episode each
    full_rebuild = metrics_scan = None
    if world == 1 and not args.no_full:
        vg.close()
        for mode in ("obs", "metrics"):
            vg2 = VecGame(geom, insts, device=local, obs_mode=1 if mode == "obs" else 0, metrics_mode=1)
            ms2, _ = _episodes(vg2, sched_p, N_NETS, 1)
            vg2.profile(True)
            t0 = N_NETS * 2
            vg2.reset()
            for t in range(N_NETS):
                vg2.step(sched_p[t0 + t])
            pr = vg2.profile_get()
            vg2.profile(False)
            if mode == "obs":
                # bytes the step moves: the observation blocks rebuilt (4 (2 + 7 n_left) cells per environment) + the scan
                step_bytes = sum(4.0 * (2 + 7 * (N_NETS - 1 - s)) * cells * ENVS_PER_GPU for s in range(N_NETS)) + 4.0 * cells * ENVS_PER_GPU * N_NETS
                step_bytes += 4.0 * (2 + 7 * N_NETS) * cells * ENVS_PER_GPU      # the episode's reset build
                full_rebuild = {"value": N_NETS * ENVS_PER_GPU / (ms2 / 1e3), "unit": "env-steps/s", "ms_per_step": ms2 / N_NETS,
                                "note": "obs_mode=1, metrics_mode=1: every stepped environment's observation is rebuilt from scratch and "
                                        "its congestion metrics are recomputed by a scan; one episode of 32 steps incl. the reset"}
                roofline["full_rebuild_step"] = {"bytes_per_step": step_bytes / N_NETS, "ms_per_step": ms2 / N_NETS,
                                                 "achieved": round(step_bytes / (ms2 / 1e3) / 1e9, 1),
                                                 "frac": round(step_bytes / (ms2 / 1e3) / 1e9 / peak, 4),
                                                 "note": "whole step (route + scan + rebuild) over the bytes of scan + rebuild"}
                ob = sum(4.0 * (2 + 7 * (N_NETS - 1 - s)) * cells * ENVS_PER_GPU for s in range(N_NETS)) + 4.0 * (2 + 7 * N_NETS) * cells * ENVS_PER_GPU
                hbm["k_obs"]["in_step"] = {"achieved": round(ob / (pr["obs"]["ms"] / 1e3) / 1e9, 1),
                                           "frac": round(ob / (pr["obs"]["ms"] / 1e3) / 1e9 / peak, 4), "launches": int(pr["obs"]["launches"]),
                                           "what": "full rebuild inside the step (obs_mode 1), CUDA events per launch over one episode"}
            else:
                mb = 4.0 * cells * ENVS_PER_GPU * pr["metrics"]["launches"]
                metrics_scan = {"value": N_NETS * ENVS_PER_GPU / (ms2 / 1e3), "unit": "env-steps/s", "ms_per_step": ms2 / N_NETS,
                                "note": "metrics_mode=1: congestion metrics recomputed by a scan every step (default: maintained by the commits)"}
                hbm["k_metrics"]["in_step"] = {"achieved": round(mb / (pr["metrics"]["ms"] / 1e3) / 1e9, 1),
                                               "frac": round(mb / (pr["metrics"]["ms"] / 1e3) / 1e9 / peak, 4),
                                               "launches": int(pr["metrics"]["launches"]), "avg_launch_us": round(1e3 * pr["metrics"]["ms"] / max(1, pr["metrics"]["launches"]), 2),
                                               "what": "the scan inside the step (metrics_mode 1): one launch over the stepped environments, CUDA events"}
            vg2.close()
            del vg2
        torch.cuda.empty_cache()

    # ---- the other configurations of BASELINE.json at their single-GPU sizes, and the real ispd18_test1 region
    legs, ispd = None, None
    if rank == 0 and world == 1 and not args.no_legs:
        try:
            vg.close()
        except Exception:
            pass
        legs = {
            "t1_7x7_512": synthetic_leg(local, "T1-7x7", 512, 32, 2, cpu_budget=0.0 if args.no_cpu else 10.0),
            "t1_1x1_1024": synthetic_leg(local, "T1-1x1", 1024, 32, 2),
            "syn1024_32": synthetic_leg(local, "SYN-1024", 32, 128, 1, obs_cap=8, steps=32),
        }
        legs["t1_7x7_512"]["config"] = "BASELINE.json configs[2]: one GPU's shard (512 of 4096 environments) of the ispd18_test1-sized grid"
        legs["t1_1x1_1024"]["config"] = "BASELINE.json configs[4]: one GPU's shard (1024 of 8192 environments), DLPack observations"
        legs["syn1024_32"]["config"] = ("BASELINE.json configs[3]: one GPU's shard (32 of 256 environments), 128 clustered nets, observation "
                                        "materialised for obstacle + order + the first 8 nets (the full one is 34 GB per environment); 32 steps")
    if rank == 0 and world == 1 and not args.no_ispd:
        ispd = ispd_leg(local, cpu=not args.no_cpu)
    # BASELINE.json configs[4] with an agent in the loop: the reference's RepresentationNetwork, batched over nets and
    # environments (xroute_env_b200/agent.py, plain PyTorch), scoring 1024 GPU environments through DLPack
    rollout = None
    if rank == 0 and world == 1 and not args.no_legs:
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "ppo_rollout.py"), "--envs", "1024", "--nets", "16",
                                  "--steps", "48", "--json"], capture_output=True, text=True, timeout=600)
            rollout = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            rollout["note"] = ("PPO-style rollout: policy (batched reference RepresentationNetwork + heads, eager PyTorch) and environment "
                               "step alternate on one stream; the split shows which of the two bounds configs[4]")
        except Exception as ex:
            rollout = {"error": repr(ex)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        oracle.build()
        v, steps, settled_s, wall = cpu_run(1, 15.0)
        cpu_baseline = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{steps} env-steps of the same workload (SYN-256, 32 nets) in {wall:.1f}s, "
                                  "oracle/xr_oracle.c obs+route+reward, single thread",
                        "cells_settled_per_s": settled_s}

    if rank == 0:
        h2d = ENVS_PER_GPU * 4
        d2h = ENVS_PER_GPU * (3 * 4 + 1 + 6 * 8)
        line = {
            "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_dev / K,
            "ms_per_step_by_rank": {"min": min(rank_ms), "mean": statistics.mean(rank_ms), "max": max(rank_ms), "all": [round(v, 4) for v in rank_ms]},
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"SYN-256 256x256x9 grid, {ENVS_PER_GPU} envs/GPU, {N_NETS} nets/env, random net "
                                   "order, back-to-back episodes (reset every 32 steps); obs+route+reward per env-step",
                       "grid": "256x256x9", "envs_per_gpu": ENVS_PER_GPU, "nets_per_env": N_NETS,
                       "l2": "working set (34 GB observations + 0.8 GB router state per GPU) >> 126 MB L2",
                       "obs_update": "in place, incremental (bit-exact vs a rebuild, tests/test_gpu_parity.py); the "
                                     "full-rebuild figure is full_obs_rebuild",
                       "engine": "frontier search (one CTA per net) + sweep kernels for wide few-pin nets (hybrid)",
                       "parallelism": f"env-shard x{world}"},
            "net_routes_per_s": value,
            "cells_relaxed_per_s": cells_relaxed_all / (ms_dev / 1e3),
            "relax_passes_per_step": relax_passes_all / (K * world),
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K,
                    "note": "xr_step with host actions + xr_step_results to pinned host (delta, done, cum); "
                            "observations stay on the GPU (DLPack) by design"},
            "long_run": {"value": LONG_EPISODES * N_NETS * ENVS_PER_GPU * world / (ms_long / 1e3), "unit": "env-steps/s",
                         "steps": LONG_EPISODES * N_NETS, "ms_total": ms_long,
                         "ms_per_step_by_rank": {"min": min(long_rank_ms), "mean": statistics.mean(long_rank_ms), "max": max(long_rank_ms)},
                         "note": "10 whole episodes end to end (host actions in, results read back every step), max over ranks"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kern,
            "profiled_leg_ms_per_step": ms_prof / K,
            "cpu_baseline": cpu_baseline,
            "full_obs_rebuild": full_rebuild,
            "metrics_scan": metrics_scan,
            "legs": legs,
            "ispd18_test1": ispd,
            "ppo_rollout": rollout,
            "route_paths": route_paths,
            "episode_stats": stats,
        }
        if legs and "x_over_cpu_all_cores" in legs["t1_7x7_512"]:
            line["x_over_cpu_all_cores_ispd18_test1_sized_grid"] = legs["t1_7x7_512"]["x_over_cpu_all_cores"]
        print(json.dumps(line), flush=True)
    try:
        vg.close()
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (cpu_baseline and the all-core arms)")
    ap.add_argument("--no-full", action="store_true", help="skip the full-observation-rebuild / metric-scan legs")
    ap.add_argument("--no-ispd", action="store_true", help="skip the real ispd18_test1 region leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the T1-7x7 / T1-1x1 / SYN-1024 legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
