"""CPU-only study for DESIGN.md section 12 item 2 (goal-directed pruning): how many cells does a connection of the bench
workload really need?  For every connection of a sample of SYN-256 episodes the oracle's converged distance field d is
compared with the best target distance B:
    need   = #{d <= B}              what any exact search must settle (the cap the sweeps already use)
    astar  = #{d + h <= B}          with h = L1 track distance (DBU) to the nearest unconnected access point -- the
                                    admissible bound of the window exit test, used as an A*-style filter
    box    = #{d + hbox <= B}       with the cheaper bound hbox = L1 distance to the bounding box of those access points
    window = cells of the net's window (access-point box + margin 8, all layers): what the window kernels relax
    python tools/analyze_pruning.py [n_envs] [margin]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import make_batch, preset_geometry
from oracle.oracle import OracleEnv

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
margin = int(sys.argv[2]) if len(sys.argv) > 2 else 8
geom = preset_geometry("SYN-256")
insts = make_batch(geom, n_envs, 32, 0)
xc, yc = geom.x_coords.astype(np.int64), geom.y_coords.astype(np.int64)
rows = []
for e, inst in enumerate(insts):
    lead, lag = OracleEnv(geom, inst), OracleEnv(geom, inst)
    order = np.random.default_rng(e).permutation(inst.net_ids)
    for net in order:
        net = int(net)
        lead.step(net)
        cells, off, cost = lead.last_paths()
        sel = inst.ap_net == net
        ap_xyz, ap_pin = inst.ap_xyz[sel], inst.ap_pin[sel]
        if len(cost) == 0:
            lag.step(net); continue
        x0, x1 = max(0, ap_xyz[:, 0].min() - margin), min(geom.X - 1, ap_xyz[:, 0].max() + margin)
        y0, y1 = max(0, ap_xyz[:, 1].min() - margin), min(geom.Y - 1, ap_xyz[:, 1].max() + margin)
        n_window = int((x1 - x0 + 1) * (y1 - y0 + 1) * geom.Z)
        src_pin = lag.src_pin(net)
        ap_ci = (ap_xyz[:, 2] * geom.Y + ap_xyz[:, 1]) * geom.X + ap_xyz[:, 0]
        connected = {int(src_pin)}
        for k in range(len(cost)):
            srcs = ap_ci[ap_pin == src_pin] if k == 0 else np.unique(cells[:off[k]])
            d = lag.distance_field(net, srcs).astype(np.int64)
            B = int(cost[k])
            tgt = ap_xyz[~np.isin(ap_pin, list(connected))]
            hx = np.abs(xc[None, :] - xc[tgt[:, 0]][:, None])            # [T, X]
            hy = np.abs(yc[None, :] - yc[tgt[:, 1]][:, None])            # [T, Y]
            h = (hx[:, None, :] + hy[:, :, None]).min(0)                 # [Y, X]
            need = int((d <= B).sum())
            astar = int((d + h[None] <= B).sum())
            bx = np.maximum(0, np.maximum(xc[tgt[:, 0]].min() - xc, xc - xc[tgt[:, 0]].max()))      # distance to the targets' box
            by = np.maximum(0, np.maximum(yc[tgt[:, 1]].min() - yc, yc - yc[tgt[:, 1]].max()))
            abox = int((d + (bx[None, :] + by[:, None])[None] <= B).sum())
            rows.append((len(set(ap_pin.tolist())), k, n_window, need, astar, off[k + 1] - off[k], abox))
            tree = set(cells[:off[k + 1]].tolist())
            connected |= {int(p) for p, c in zip(ap_pin, ap_ci) if int(c) in tree}
        lag.step(net)
r = np.array(rows, np.int64)
print(f"{len(r)} connections of {n_envs} SYN-256 episodes (32 nets each), grid {geom.cells} cells")
for name, m in (("all", np.ones(len(r), bool)), ("first connection", r[:, 1] == 0), ("later connections", r[:, 1] > 0),
                ("nets with >= 8 pins", r[:, 0] >= 8), (">= 8 pins, later connections", (r[:, 0] >= 8) & (r[:, 1] > 0))):
    if m.sum() == 0:
        continue
    w, need, astar, plen, abox = r[m, 2].sum(), r[m, 3].sum(), r[m, 4].sum(), r[m, 5].sum(), r[m, 6].sum()
    print(f"  {name:32s} n={int(m.sum()):4d}  window cells {w:>11d}  d<=B {need:>10d} ({need / w:.3f} of window)  "
          f"d+h<=B {astar:>9d} ({astar / w:.4f} of window, {astar / max(need, 1):.3f} of d<=B)  box bound {abox / w:.3f} of window  path cells {plen}")
