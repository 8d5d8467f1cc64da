"""Debug: which window-engine settings reproduce the oracle on a small batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import VecGame, make_batch, ispd18_geometry
from oracle.oracle import OracleEnv

def run(shape, n_env, n_nets, seed, **kw):
    geom = ispd18_geometry(*shape)
    insts = make_batch(geom, n_env, n_nets, seed=seed, p_obstacle=kw.pop("p_obstacle", 0.1))
    vg = VecGame(geom, insts, device=0, **kw)
    vg.reset()
    orcs = [OracleEnv(geom, i) for i in insts]
    rng = np.random.default_rng(1)
    orders = [list(rng.permutation(i.net_ids)) for i in insts]
    bad = []
    try:
        for t in range(n_nets):
            acts = np.array([int(o[t]) for o in orders], np.int32)
            vg.step(acts)
            for e, o in enumerate(orcs):
                o.step(int(acts[e]))
                oc, oo, ocost = o.last_paths(); gc, go, gcost = vg.paths(e)
                if not (np.array_equal(oc, gc) and np.array_equal(ocost, gcost)):
                    bad.append((t, e, len(ocost), ocost.tolist()[:4], gcost.tolist()[:4]))
    except Exception as ex:
        bad.append(("EXC", str(ex)[:80], t))
    vg.close()
    return bad

for shape in [(25, 26, 9), (48, 44, 9), (90, 90, 9)]:
    for kw in [dict(min_cluster=1), dict(min_cluster=2), dict(min_cluster=4), dict(min_cluster=8), dict(min_cluster=16), dict(window_margin=-1)]:
        bad = run(shape, 4, 8, 100 + shape[0], **dict(kw))
        print(shape, kw, "OK" if not bad else f"BAD x{len(bad)}: {bad[:3]}", flush=True)
