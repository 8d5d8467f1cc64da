"""CPU emulation of the frontier kernel (csrc/xr_frontier.cu) as it is built: per connection a fresh search from the
whole tree on a sparse field, open list of (cell, d, f) entries, per ROUND every entry with f <= fmin + delta is
expanded, an expansion relaxes a RAY of up to L cells in each of the six directions (L per direction class), the
heuristic is the L1 track distance to the nearest unconnected pin's access-point BOX.  Checks target, cost and
canonical path against the oracle and reports rounds / expansions / relaxations / open-list size.
    python tools/simulate_frontier2.py [preset] [n_envs] [delta] [Lpref] [Lnonpref] [Lvia]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import make_batch, preset_geometry
from oracle.oracle import OracleEnv

preset = sys.argv[1] if len(sys.argv) > 1 else "SYN-256"
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
delta = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
LP = int(sys.argv[4]) if len(sys.argv) > 4 else 4
LN = int(sys.argv[5]) if len(sys.argv) > 5 else 1
LV = int(sys.argv[6]) if len(sys.argv) > 6 else 1
STOP = int(sys.argv[7]) if len(sys.argv) > 7 else 1   # a ray ends at the first cell it does not lower
geom = preset_geometry(preset)
X, Y, Z = geom.X, geom.Y, geom.Z
xc, yc = [int(v) for v in geom.x_coords], [int(v) for v in geom.y_coords]
DELTA = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
INF = 1 << 60
ldir = [int(v) for v in geom.layer_dir]
pitch = [int(v) for v in geom.layer_pitch]
minw = [int(v) for v in geom.layer_min_width]


def search(cflag, tree, tpins):
    """tpins: list of (pin, [cells]) of the unconnected pins."""
    boxes = []
    tset = set()
    for p, cells in tpins:
        xs = [xc[c[0]] for c in cells]; ys = [yc[c[1]] for c in cells]
        boxes.append((min(xs), max(xs), min(ys), max(ys)))
        tset |= set(cells)
    def h(c):
        px, py = xc[c[0]], yc[c[1]]
        best = INF
        for (x0, x1, y0, y1) in boxes:
            v = max(x0 - px, 0, px - x1) + max(y0 - py, 0, py - y1)
            if v < best:
                best = v
        return best
    def w(p, c):
        f = int(cflag[c[2], c[1], c[0]])
        mult = 1 + geom.drc_cost * (f & 1) + geom.fixed_shape_cost * ((f >> 1) & 1)
        pen = geom.block_cost * minw[c[2]] * 20 * ((f >> 2) & 1)
        if p[2] != c[2]:
            return geom.via_cost * pitch[max(p[2], c[2])] * mult + pen
        axis = 0 if p[0] != c[0] else 1
        length = abs(xc[c[0]] - xc[p[0]]) + abs(yc[c[1]] - yc[p[1]])
        return length * (mult + geom.grid_cost * (axis != ldir[c[2]])) + pen
    d = {c: 0 for c in tree}
    B = 0 if tset & set(tree) else INF
    lst = [(c, 0, h(c)) for c in tree]
    rounds = expansions = relax = widest = maxopen = stale = 0
    while True:
        lst = [e for e in lst if e[2] <= B]
        if not lst:
            break
        maxopen = max(maxopen, len(lst))
        fmin = min(e[2] for e in lst)
        thr = fmin + delta
        nxt, batch = [], []
        for e in lst:
            (batch if e[2] <= thr else nxt).append(e)
        rounds += 1; widest = max(widest, len(batch))
        for (c, dc, f) in batch:
            if d[c] < dc:
                stale += 1
                continue
            expansions += 1
            for k in range(6):
                if k >= 4:
                    L = LV
                else:
                    L = LP if (k >> 1) == ldir[c[2]] else LN
                u, nd = c, dc
                for s in range(L):
                    v = (u[0] + DELTA[k][0], u[1] + DELTA[k][1], u[2] + DELTA[k][2])
                    if not (0 <= v[0] < X and 0 <= v[1] < Y and 0 <= v[2] < Z):
                        break
                    nd = nd + w(u, v)
                    if nd > B:
                        break
                    if STOP and nd >= d.get(v, INF):
                        break
                    relax += 1
                    if nd < d.get(v, INF):
                        d[v] = nd
                        if v in tset and nd < B:
                            B = nd
                        fv = nd + h(v)
                        if fv <= B:
                            nxt.append((v, nd, fv))
                    u = v
        lst = nxt
    return d, w, rounds, expansions, relax, widest, maxopen, stale


def walk(d, w, t):
    c, last, path = t, None, [t]
    while d[c] != 0:
        for k in ([last] if last is not None else []) + list(range(6)):
            p = (c[0] - DELTA[k][0], c[1] - DELTA[k][1], c[2] - DELTA[k][2])
            if p in d and d[p] + w(p, c) == d[c]:
                break
        else:
            raise AssertionError("no predecessor")
        c, last = p, k
        path.append(c)
    return path


kw = dict(hot_spots=16, p_obstacle=0.10) if preset == "SYN-1024" else {}
n_nets = 128 if preset == "SYN-1024" else 32
insts = make_batch(geom, n_envs, n_nets, 0, **kw)
rows = []
for e, inst in enumerate(insts):
    lead, lag = OracleEnv(geom, inst), OracleEnv(geom, inst)
    apnet = np.zeros((Z, Y, X), np.int64); apnet[inst.ap_xyz[:, 2], inst.ap_xyz[:, 1], inst.ap_xyz[:, 0]] = inst.ap_net
    blk = np.zeros((Z, Y, X), np.uint8)
    if len(inst.block_xyz):
        blk[inst.block_xyz[:, 2], inst.block_xyz[:, 1], inst.block_xyz[:, 0]] = 1
    order = np.random.default_rng(e).permutation(inst.net_ids)
    if preset == "SYN-1024":
        order = order[:12]
    for net in order:
        net = int(net)
        lead.step(net)
        cells, off, cost = lead.last_paths()
        sel = inst.ap_net == net
        aps = [(int(p), tuple(int(v) for v in xyz)) for p, xyz in zip(inst.ap_pin[sel], inst.ap_xyz[sel])]
        if len(cost):
            usage = lag.state()[0]
            cflag = ((usage > 0).astype(np.uint8) | (((apnet != 0) & (apnet != net)).astype(np.uint8) << 1) | (blk << 2))
            src_pin = int(lag.src_pin(net))
            tree, connected = [xyz for p, xyz in aps if p == src_pin], {src_pin}
            for k in range(len(cost)):
                pins = sorted(set(p for p, _ in aps if p not in connected))
                tpins = [(p, [xyz for q, xyz in aps if q == p]) for p in pins]
                targets = [xyz for p, xyz in aps if p not in connected]
                d, w, rounds, expansions, relax, widest, maxopen, stale = search(cflag, tree, tpins)
                t = min((d.get(c, INF), (c[2] * Y + c[1]) * X + c[0], c) for c in targets)
                path = walk(d, w, t[2])
                want = cells[off[k]:off[k + 1]].tolist()
                assert t[0] == int(cost[k]) and [(z * Y + y) * X + x for (x, y, z) in path] == want, (e, net, k)
                rows.append((len(set(p for p, _ in aps)), k, rounds, expansions, relax, widest, maxopen, len(path), len(d), stale))
                tree = path if k == 0 else tree + path
                on = set(tree)
                connected |= {p for p, xyz in aps if xyz in on}
        lag.step(net)
r = np.array(rows, np.int64)
print(f"{preset} delta {delta} L {LP}/{LN}/{LV}: {len(r)} connections, every target / cost / path == oracle")
for name, m in (("all", np.ones(len(r), bool)), ("first, 2-3 pin nets", (r[:, 1] == 0) & (r[:, 0] <= 3)),
                ("first, 4-7 pin nets", (r[:, 1] == 0) & (r[:, 0] >= 4) & (r[:, 0] <= 7)),
                ("first, >= 8 pin nets", (r[:, 1] == 0) & (r[:, 0] >= 8)), ("later connections", r[:, 1] > 0),
                (">= 8 pins, later connections", (r[:, 0] >= 8) & (r[:, 1] > 0))):
    if m.sum():
        q = r[m]
        print(f"   {name:30s} n={len(q):4d}  rounds mean {q[:, 2].mean():7.1f} p90 {np.percentile(q[:, 2], 90):6.0f} max {q[:, 2].max():5d} | "
              f"expans mean {q[:, 3].mean():8.1f} max {q[:, 3].max():6d} | relax mean {q[:, 4].mean():9.1f} max {q[:, 4].max():7d} | widest mean {q[:, 5].mean():6.1f} max {q[:, 5].max():5d} | "
              f"open max {q[:, 6].max():6d} | touched mean {q[:, 8].mean():8.1f} max {q[:, 8].max():6d} | stale {q[:, 9].sum()}")
