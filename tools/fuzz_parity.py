"""Randomised parity campaign: random grid shapes (uniform and non-uniform pitch), obstacle densities, net counts and
engine settings; every connection cost, path, metric and the final occupancy against the CPU oracle.
    python tools/fuzz_parity.py [seconds] [seed] [ties]
`ties`: every pitch, via cost and penalty is a small multiple of 100 (an x step, a wrong-way y step and a via often weigh
the same), layer directions are random and the cost constants vary -- almost every cell then has several equal-cost
predecessors, which is what exercises the canonical target and backtrace rules (the one mismatch this campaign ever
found needed x pitch == 3 x y pitch next to a source)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import VecGame, make_batch, ispd18_geometry
from oracle.oracle import OracleEnv

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ties = len(sys.argv) > 3 and sys.argv[3] == "ties"
t0, n_cfg, n_steps, bad = time.time(), 0, 0, []
while time.time() - t0 < budget and not bad:
    X, Y, Z = int(rng.integers(8, 130)), int(rng.integers(8, 130)), int(rng.integers(2, 10))
    geom = ispd18_geometry(X, Y, Z)
    if ties:
        px, py = rng.choice([100, 200, 300], 2)
        geom.x_coords = (np.cumsum(rng.choice([1, 1, 1, 2, 3], X)) * px).astype(np.int32) if rng.random() < 0.3 else (px * np.arange(X)).astype(np.int32)
        geom.y_coords = (np.cumsum(rng.choice([1, 1, 1, 2, 3], Y)) * py).astype(np.int32) if rng.random() < 0.3 else (py * np.arange(Y)).astype(np.int32)
        geom.layer_dir = rng.integers(0, 2, Z).astype(np.uint8)
        geom.layer_pitch = rng.choice([25, 50, 75, 100], Z).astype(np.int32)
        geom.layer_min_width = rng.choice([5, 10, 20], Z).astype(np.int32)
        geom.via_cost, geom.grid_cost = int(rng.choice([1, 2, 4])), int(rng.choice([0, 1, 2]))
        geom.drc_cost, geom.fixed_shape_cost, geom.block_cost = int(rng.choice([1, 2, 8])), int(rng.choice([1, 2, 8])), int(rng.choice([1, 5, 32]))
    elif rng.random() < 0.4:
        geom.x_coords = np.cumsum(rng.integers(60, 900, X)).astype(np.int32)
        geom.y_coords = np.cumsum(rng.integers(60, 900, Y)).astype(np.int32)
    n_env, n_nets = int(rng.integers(2, 7)), int(rng.integers(2, 11))
    iseed, pob = int(rng.integers(1 << 30)), float(rng.choice([0.0, 0.1, 0.25, 0.4]))
    uniform = len(set(np.diff(geom.x_coords).tolist())) == 1
    try:
        insts = make_batch(geom, n_env, n_nets, seed=iseed, p_obstacle=pob)
    except RuntimeError:
        continue
    kw = dict(window_margin=int(rng.choice([-1, 0, 0, 0, 1, 4, 30])), min_cluster=int(rng.choice([0, 0, 1, 2, 4, 8, 16])),
              obs_mode=int(rng.choice([0, 0, 1])))
    dual = rng.random() < 0.5
    # engine 0 (default): the frontier search with random ray length / bucket width / block size / list capacity and the
    # hybrid threshold anywhere from "every net" to "no net" on the sweep kernels; engine 1: the sweep engines alone
    kw["engine"] = int(rng.random() < 0.3)
    kw["metrics_mode"] = int(rng.random() < 0.3)
    if kw["engine"] == 0 and rng.random() < 0.35:         # the optional cost terms (frontier engine only), with synthetic guides
        kw["guide_cost"], kw["halo"] = int(rng.choice([0, 1, 4])), int(rng.choice([0, 1, 2]))
        for inst in insts:
            boxes = []
            for n in inst.net_ids:
                xy = inst.ap_xyz[inst.ap_net == n]
                m = int(rng.integers(0, 3))
                for z in range(0, Z, int(rng.integers(1, 3))):
                    if rng.random() < 0.8:
                        boxes.append((n, xy[:, 0].min() - m, xy[:, 0].max() + m, xy[:, 1].min() - m, xy[:, 1].max() + m, z))
            inst.guides = np.array(boxes, np.int32).reshape(-1, 6)
    knobs = {"XR_FR_RAY": str(int(rng.integers(1, 9))), "XR_FR_DELTA": str(int(rng.choice([0, 100, 400, 1200, 100000]))),
             "XR_FR_DMAX": str(int(rng.choice([1, 4, 16]))), "XR_FR_THREADS": str(int(rng.choice([64, 256, 1024]))),
             "XR_HYBRID_AREA": str(int(rng.choice([0, 30, 400, 4000]))), "XR_HYBRID_PINS": str(int(rng.choice([2, 3, 30]))),
             "XR_FR_PARK": str(int(rng.choice([0, 1, 16, 256, 4096]))), "XR_FR_BAND": str(int(rng.choice([1, 2, 4, 64])))}
    if rng.random() < 0.4:
        knobs["XR_WIN_FIT_CAP"] = str(int(rng.choice([3000, 9000, 30000])))
    else:
        os.environ.pop("XR_WIN_FIT_CAP", None)
    if rng.random() < 0.3:
        knobs["XR_FR_CAP"] = "64"
    else:
        os.environ.pop("XR_FR_CAP", None)
    os.environ.update(knobs)
    for k, v in (("XR_DUAL_PINS", "2" if dual else "8"), ("XR_DUAL_MINC", str(int(rng.choice([2, 4, 8]))) if dual else "8")):
        os.environ[k] = v
    vg = VecGame(geom, insts, device=0, **kw)
    vg.reset()
    orcs = [OracleEnv(geom, i, guide_cost=kw.get("guide_cost", 0), halo=kw.get("halo", 0)) for i in insts]
    orders = [list(rng.permutation(i.net_ids)) for i in insts]
    for t in range(max(len(o) for o in orders)):
        acts = np.array([int(o[t]) if t < len(o) else 0 for o in orders], np.int32)
        try:
            vg.step(acts)
        except Exception as ex:                 # keep the configuration of a failing step for tools/repro_fuzz2.py
            import pickle
            os.makedirs("gpurun_out", exist_ok=True)
            with open("gpurun_out/fuzz_fail.pkl", "wb") as fh:
                pickle.dump(dict(geom=geom, insts=insts, kw=kw, knobs=knobs, dual=dual, orders=orders, t=t,
                                 env={k: os.environ.get(k) for k in ("XR_DUAL_PINS", "XR_DUAL_MINC")}), fh)
            bad.append(dict(shape=(X, Y, Z), kw=kw, knobs=knobs, t=t, error=str(ex)))
            break
        delta, done, cum = vg.results_host()
        for e, o in enumerate(orcs):
            if acts[e] == 0:
                continue
            m = o.step(int(acts[e]))
            oc, oo, ocost = o.last_paths(); gc, go, gcost = vg.paths(e)
            ok = (np.array_equal(oc, gc) and np.array_equal(ocost, gcost) and
                  [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"], m["overflow"]])
            if ok and (t % 3 == 0 or m["done"]):
                ok = np.array_equal(vg.obs_host(e).numpy(), o.obs())
            if not ok:
                pins = len(set(insts[e].ap_pin[insts[e].ap_net == acts[e]].tolist()))
                bad.append(dict(shape=(X, Y, Z), uniform=uniform, n_env=n_env, n_nets=n_nets, iseed=iseed, pob=pob, kw=kw, dual=dual, knobs=knobs,
                                minc=os.environ["XR_DUAL_MINC"], t=t, e=e, net=int(acts[e]), pins=pins,
                                cost_o=ocost.tolist()[:6], cost_g=gcost.tolist()[:6], path_eq=bool(np.array_equal(oc, gc)),
                                cum_g=[int(v) for v in cum[e]], cum_o=[m["violation"], m["wirelength"], m["via"]],
                                rc=vg.route_counters()))
            n_steps += 1
    for e, o in enumerate(orcs):
        if bad:
            break
        if not (np.array_equal(vg.state(e)[0], o.state()[0]) and np.array_equal(vg.state(e)[1], o.state()[1])):
            bad.append(dict(shape=(X, Y, Z), kw=kw, dual=dual, what="state", e=e))
    vg.close()
    n_cfg += 1
print(f"fuzz: {n_cfg} configurations, {n_steps} env-steps in {time.time() - t0:.0f}s:", "ALL BIT-EXACT" if not bad else f"MISMATCH {bad[:3]}")
sys.exit(1 if bad else 0)
