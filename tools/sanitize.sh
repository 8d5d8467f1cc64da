# compute-sanitizer over a small episode through every routing engine: frontier search (default, with list spill, with the
# optional guide / halo terms), band kernel, dual kernel, full-grid sweeps, forced window escapes
#   bash tools/sanitize.sh [memcheck|racecheck|synccheck]
set -e
TOOL=${1:-memcheck}
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from xroute_env_b200 import VecGame, make_batch, ispd18_geometry
geom = ispd18_geometry(40, 36, 9)
insts = make_batch(geom, 3, 6, seed=7)
for inst in insts:
    boxes = []
    for n in inst.net_ids:
        xy = inst.ap_xyz[inst.ap_net == n]
        for z in (0, 2):
            boxes.append((n, xy[:, 0].min(), xy[:, 0].max(), xy[:, 1].min(), xy[:, 1].max(), z))
    inst.guides = np.array(boxes, np.int32)
for kw in (dict(), dict(guide_cost=1, halo=1), dict(metrics_mode=1, obs_mode=1), dict(engine=1), dict(engine=1, min_cluster=2),
           dict(engine=1, window_margin=-1), dict(engine=1, window_margin=1)):
    vg = VecGame(geom, insts, device=0, **kw)
    vg.reset()
    rng = np.random.default_rng(0)
    orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)
    for t in range(6):
        vg.step_async(orders[t]); vg.step_wait()
    vg.results_host()
    print(kw, vg.route_counters(), flush=True)
    vg.close()
PY
run() { compute-sanitizer --tool $TOOL --error-exitcode 9 python /tmp/san.py 2>&1 | grep -A1 "Error: Race\|ERROR SUMMARY\|RACECHECK SUMMARY\|Invalid\|Hazard" | cut -c1-260 | sort | uniq -c | sort -rn | head -40; }
echo "== default knobs"; run
echo "== frontier lists forced through their global spill, 64-thread blocks, short rays"; XR_FR_CAP=64 XR_FR_THREADS=64 XR_FR_RAY=3 run
echo "== every sweep-engine net through the dual cyclic layout kernel"; XR_DUAL_PINS=2 XR_DUAL_MINC=2 run
echo "== wide variant of the frontier kernel: far list from the first open entry on"; XR_FR_PARK=1 XR_FR_BAND=1 run
echo "== hybrid: wide nets on the sweep kernels beside the frontier kernel"; XR_HYBRID_AREA=30 XR_HYBRID_PINS=30 run
