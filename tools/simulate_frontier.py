"""CPU simulation of the goal-directed frontier search proposed in DESIGN.md section 12 item 0, on the bench workload,
to size the GPU kernel before writing it: per connection a label-correcting search from the whole tree, open cells
expanded in parallel ROUNDS -- every open cell with f = d + h <= (smallest open f) + delta is expanded in the same round
-- until no open cell has f <= B (best tentative target distance).  Reports rounds (the serial depth a CTA / warp would
pay), expansions (work), the widest round, and checks target, cost and canonical path against the oracle.
    python tools/simulate_frontier.py [n_envs] [delta_dbu ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xroute_env_b200 import make_batch, preset_geometry
from oracle.oracle import OracleEnv

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
deltas = [int(v) for v in sys.argv[2:]] or [0, 400, 1200]
geom = preset_geometry("SYN-256")
X, Y, Z = geom.X, geom.Y, geom.Z
xc, yc = [int(v) for v in geom.x_coords], [int(v) for v in geom.y_coords]
DELTA = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
INF = 1 << 60


def search(cflag, tree, targets, delta):
    """Returns (d dict, rounds, expansions, widest)."""
    tx = np.array([xc[t[0]] for t in targets], np.int64); ty = np.array([yc[t[1]] for t in targets], np.int64)
    hcache = {}
    def h(c):
        v = hcache.get((c[0], c[1]))
        if v is None:
            v = int((np.abs(tx - xc[c[0]]) + np.abs(ty - yc[c[1]])).min()); hcache[(c[0], c[1])] = v
        return v
    def w(p, c):
        f = int(cflag[c[2], c[1], c[0]])
        mult = 1 + geom.drc_cost * (f & 1) + geom.fixed_shape_cost * ((f >> 1) & 1)
        pen = geom.block_cost * int(geom.layer_min_width[c[2]]) * 20 * ((f >> 2) & 1)
        if p[2] != c[2]:
            return geom.via_cost * int(geom.layer_pitch[max(p[2], c[2])]) * mult + pen
        axis = 0 if p[0] != c[0] else 1
        length = abs(xc[c[0]] - xc[p[0]]) + abs(yc[c[1]] - yc[p[1]])
        return length * (mult + geom.grid_cost * (axis != int(geom.layer_dir[c[2]]))) + pen
    tset = set(targets)
    d = {c: 0 for c in tree}
    open_ = {c: h(c) for c in tree}
    best = 0 if tset & set(tree) else INF
    rounds = expansions = widest = 0
    while open_:
        fmin = min(open_.values())
        if fmin > best:
            break
        batch = [c for c, f in open_.items() if f <= fmin + delta and f <= best]
        for c in batch:
            del open_[c]
        rounds += 1; expansions += len(batch); widest = max(widest, len(batch))
        for c in batch:
            dc = d[c]
            for k in range(6):
                v = (c[0] + DELTA[k][0], c[1] + DELTA[k][1], c[2] + DELTA[k][2])
                if not (0 <= v[0] < X and 0 <= v[1] < Y and 0 <= v[2] < Z):
                    continue
                nd = dc + w(c, v)
                if nd < d.get(v, INF):
                    d[v] = nd; open_[v] = nd + h(v)
                    if v in tset and nd < best:
                        best = nd
    return d, w, rounds, expansions, widest


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


insts = make_batch(geom, n_envs, 32, 0)
for delta in deltas:
    rows = []
    for e, inst in enumerate(insts):
        lead, lag = OracleEnv(geom, inst), OracleEnv(geom, inst)
        apnet = np.zeros((Z, Y, X), np.int64); apnet[inst.ap_xyz[:, 2], inst.ap_xyz[:, 1], inst.ap_xyz[:, 0]] = inst.ap_net
        blk = np.zeros((Z, Y, X), np.uint8)
        if len(inst.block_xyz):
            blk[inst.block_xyz[:, 2], inst.block_xyz[:, 1], inst.block_xyz[:, 0]] = 1
        for net in np.random.default_rng(e).permutation(inst.net_ids):
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
                    targets = [xyz for p, xyz in aps if p not in connected]
                    d, w, rounds, expansions, widest = search(cflag, tree, targets, delta)
                    t = min((d.get(c, INF), (c[2] * Y + c[1]) * X + c[0], c) for c in targets)
                    path = walk(d, w, t[2])
                    want = cells[off[k]:off[k + 1]].tolist()
                    assert t[0] == int(cost[k]) and [(z * Y + y) * X + x for (x, y, z) in path] == want, (e, net, k)
                    rows.append((len(set(p for p, _ in aps)), k, rounds, expansions, widest, len(path), len(d)))
                    tree = path if k == 0 else tree + path
                    on = set(tree)
                    connected |= {p for p, xyz in aps if xyz in on}
            lag.step(net)
    r = np.array(rows, np.int64)
    print(f"delta {delta} DBU: {len(r)} connections, every target / cost / path == oracle")
    for name, m in (("all", np.ones(len(r), bool)), ("first connections", r[:, 1] == 0), ("first, 2-3 pin nets", (r[:, 1] == 0) & (r[:, 0] <= 3)),
                    ("first, >= 8 pin nets", (r[:, 1] == 0) & (r[:, 0] >= 8)), ("later connections", r[:, 1] > 0),
                    (">= 8 pins, later connections", (r[:, 0] >= 8) & (r[:, 1] > 0))):
        if m.sum():
            q = r[m]
            print(f"   {name:30s} n={len(q):4d}  rounds mean {q[:, 2].mean():7.1f} p90 {np.percentile(q[:, 2], 90):6.0f} max {q[:, 2].max():5d} | "
                  f"expansions mean {q[:, 3].mean():8.1f} max {q[:, 3].max():6d} | widest round mean {q[:, 4].mean():6.1f} max {q[:, 4].max():5d} | "
                  f"cells touched mean {q[:, 6].mean():8.1f} p99 {np.percentile(q[:, 6], 99):7.0f} max {q[:, 6].max():6d} | path cells mean {q[:, 5].mean():5.1f}")
