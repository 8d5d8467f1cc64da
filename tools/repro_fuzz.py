"""Reproduce the fuzz mismatch: shape (108,110,9) non-uniform, iseed 1063468248, env 1, net 2 (17 pins)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import VecGame, make_batch, ispd18_geometry
from oracle.oracle import OracleEnv

# regenerate the geometry exactly as the fuzz did: replay its RNG stream up to configuration 762
rng = np.random.default_rng(7)
target = None
n = 0
while target is None:
    X, Y, Z = int(rng.integers(8, 130)), int(rng.integers(8, 130)), int(rng.integers(2, 10))
    geom = ispd18_geometry(X, Y, Z)
    if rng.random() < 0.4:
        geom.x_coords = np.cumsum(rng.integers(60, 900, X)).astype(np.int32)
        geom.y_coords = np.cumsum(rng.integers(60, 900, Y)).astype(np.int32)
    n_env, n_nets = int(rng.integers(2, 7)), int(rng.integers(2, 11))
    iseed, pob = int(rng.integers(1 << 30)), float(rng.choice([0.0, 0.1, 0.25, 0.4]))
    try:
        insts = make_batch(geom, n_env, n_nets, seed=iseed, p_obstacle=pob)
    except RuntimeError:
        continue
    kw = dict(window_margin=int(rng.choice([-1, 0, 0, 0, 1, 4, 30])), min_cluster=int(rng.choice([0, 0, 1, 2, 4, 8, 16])), obs_mode=int(rng.choice([0, 0, 1])))
    dual = rng.random() < 0.5
    minc = str(int(rng.choice([2, 4, 8]))) if dual else "8"
    orders = [list(rng.permutation(i.net_ids)) for i in insts]
    if iseed == 1063468248:
        target = (geom, insts, kw, orders)
geom, insts, kw, orders = target
print("geometry", geom.X, geom.Y, geom.Z, "orders", orders)
inst = insts[1]
net = int(orders[1][0])
orc = OracleEnv(geom, inst)
orc.step(net)
oc, oo, ocost = orc.last_paths()
def ci2xyz(c): return (int(c % geom.X), int((c // geom.X) % geom.Y), int(c // (geom.X * geom.Y)))
for name, env_knobs, kws in [("band c2", {"XR_DUAL_PINS": "0"}, dict(min_cluster=2)), ("dual c8", {"XR_DUAL_PINS": "2", "XR_DUAL_MINC": "8"}, dict(min_cluster=2)),
                             ("dual c2", {"XR_DUAL_PINS": "2", "XR_DUAL_MINC": "2"}, dict(min_cluster=2)), ("global", {}, dict(window_margin=-1))]:
    os.environ.update(env_knobs)
    vg = VecGame(geom, [inst], device=0, **kws)
    vg.reset()
    vg.step(np.array([net], np.int32))
    gc, go, gcost = vg.paths(0)
    same = np.array_equal(oc, gc)
    msg = f"{name}: costs equal {np.array_equal(ocost, gcost)}, paths equal {same}"
    if not same:
        for k in range(len(ocost)):
            a, b = oc[oo[k]:oo[k + 1]], gc[go[k]:go[k + 1]]
            if not np.array_equal(a, b):
                j = next(i for i in range(min(len(a), len(b))) if a[i] != b[i]) if len(a) and len(b) else 0
                msg += f"\n   first differing connection {k} cost {ocost[k]} len {len(a)} vs {len(b)}; diverge at step {j}: oracle {[ci2xyz(c) for c in a[max(0,j-2):j+4]]} gpu {[ci2xyz(c) for c in b[max(0,j-2):j+4]]}"
                break
    print(msg, flush=True)
    vg.close()
