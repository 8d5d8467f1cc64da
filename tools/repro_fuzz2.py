"""Replay the configuration tools/fuzz_parity.py saved on a failing step (gpurun_out/fuzz_fail.pkl): which engine / knob
set fails, on which environment and net, and how its paths differ from the oracle's."""
import os, pickle, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import VecGame
from oracle.oracle import OracleEnv

cfg = pickle.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fuzz_fail.pkl", "rb"))
geom, insts, kw, knobs, orders, tfail = cfg["geom"], cfg["insts"], cfg["kw"], cfg["knobs"], cfg["orders"], cfg["t"]
print("shape", geom.X, geom.Y, geom.Z, "kw", kw, "knobs", knobs, "env", cfg["env"], "failing step", tfail)
print("x pitch set", sorted(set(np.diff(geom.x_coords).tolist()))[:6], "y", sorted(set(np.diff(geom.y_coords).tolist()))[:6],
      "dir", geom.layer_dir.tolist(), "pitch", geom.layer_pitch.tolist(), "costs", geom.via_cost, geom.grid_cost, geom.drc_cost, geom.fixed_shape_cost, geom.block_cost)
variants = [("as fuzzed", dict(kw), dict(knobs, **{k: v for k, v in cfg["env"].items() if v})),
            ("engine 0, default knobs", dict(kw, engine=0), {}),
            ("engine 0, no hybrid", dict(kw, engine=0), {"XR_HYBRID_AREA": "0"}),
            ("engine 1", dict(kw, engine=1), {k: v for k, v in cfg["env"].items() if v})]
for name, k2, env in variants:
    for k in list(os.environ):
        if k.startswith("XR_"):
            del os.environ[k]
    os.environ.update(env)
    vg = VecGame(geom, insts, device=0, **k2)
    vg.reset()
    orcs = [OracleEnv(geom, i) for i in insts]
    status = "ok"
    for t in range(tfail + 1):
        acts = np.array([int(o[t]) if t < len(o) else 0 for o in orders], np.int32)
        for e in range(len(insts)):                      # one environment at a time: isolates the failing net
            one = np.zeros_like(acts); one[e] = acts[e]
            if one[e] == 0:
                continue
            try:
                vg.step(one)
            except Exception as ex:
                pins = len(set(insts[e].ap_pin[insts[e].ap_net == acts[e]].tolist()))
                status = f"step {t} env {e} net {acts[e]} ({pins} pins): {ex}; route counters {vg.route_counters()}"
                break
            m = orcs[e].step(int(acts[e]))
            oc, oo, ocost = orcs[e].last_paths(); gc, go, gcost = vg.paths(e)
            if not (np.array_equal(oc, gc) and np.array_equal(ocost, gcost)):
                status = f"step {t} env {e} net {acts[e]}: MISMATCH costs {ocost.tolist()[:5]} vs {gcost.tolist()[:5]}"
                break
        if status != "ok":
            break
    print(f"{name}: {status}", flush=True)
    try:
        vg.close()
    except Exception:
        pass
