# compute-sanitizer over a small episode through every routing engine (band kernel, dual kernel, full grid)
#   bash tools/sanitize.sh [memcheck|racecheck]
set -e
TOOL=${1:-memcheck}
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from xroute_env_b200 import VecGame, make_batch, ispd18_geometry
geom = ispd18_geometry(40, 36, 9)
insts = make_batch(geom, 3, 6, seed=7)
for kw in (dict(), dict(min_cluster=2), dict(window_margin=-1), dict(window_margin=1)):
    vg = VecGame(geom, insts, device=0, **kw)
    vg.reset()
    rng = np.random.default_rng(0)
    orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)
    for t in range(6):
        vg.step(orders[t])
    vg.results_host()
    print(kw, vg.route_counters(), flush=True)
    vg.close()
PY
compute-sanitizer --tool $TOOL --error-exitcode 9 python /tmp/san.py 2>&1 | grep -A1 "Error: Race\|ERROR SUMMARY\|RACECHECK SUMMARY" | cut -c1-260 | sort | uniq -c | sort -rn | head -40
XR_DUAL_PINS=2 XR_DUAL_MINC=2 compute-sanitizer --tool $TOOL --error-exitcode 9 python /tmp/san.py 2>&1 | grep -A1 "Error: Race\|ERROR SUMMARY\|RACECHECK SUMMARY" | cut -c1-260 | sort | uniq -c | sort -rn | head -40
