"""Shared helpers for the parity tests."""
import os

import numpy as np

from xroute_env_b200.instances import Instance, ispd18_geometry

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rows_to_oracle(rows, dims, routed=(), infer_netlist=None):
    """Replay a golden node stream (rows of x,y,z,used,Net,Pin; tests/golden/make_golden.py)
    through the CPU oracle: blockages / APs become the instance, `used` the occupancy."""
    from oracle.oracle import OracleEnv
    X, Y, Z = (int(v) for v in dims)
    geom = ispd18_geometry(X, Y, Z)
    blk = rows[rows[:, 4] == -1][:, :3]
    ap = rows[rows[:, 4] >= 1]
    inst = Instance(block_xyz=np.ascontiguousarray(blk, np.int32).reshape(-1, 3),
                    ap_net=np.ascontiguousarray(ap[:, 4], np.int32),
                    ap_pin=np.ascontiguousarray(ap[:, 5], np.int32),
                    ap_xyz=np.ascontiguousarray(ap[:, :3], np.int32).reshape(-1, 3))
    env = OracleEnv(geom, inst)
    usage = np.zeros((Z, Y, X), np.uint8)
    for x, y, z, used, net, pin in rows:
        if used and net != -1:
            usage[z, y, x] = 1
    env.set_usage(usage)
    nets = sorted(set(int(v) for v in ap[:, 4]))
    if infer_netlist is not None:
        for n in nets:
            env.set_routed(n, n not in set(int(v) for v in infer_netlist))
    else:
        for n in routed:
            if n in nets:
                env.set_routed(int(n), True)
    return env, geom, inst


def rows_to_data(rows, dims, cum, netlist):
    nodes = [[[int(r[0]), int(r[1]), int(r[2])], [0, 0, int(r[2])], [int(r[3]), int(r[4]), int(r[5])]] for r in rows]
    return [[int(v) for v in dims], nodes, list(cum), [int(v) for v in netlist]]


def golden_obs_cases():
    z = np.load(os.path.join(GOLD, "obs_cases.npz"))
    for i in range(int(z["n_cases"][0])):
        k = f"c{i}"
        yield dict(idx=i, dims=z[k + "_dims"], rows=z[k + "_rows"], infer=bool(z[k + "_mode"][0]),
                   routed=[int(v) for v in z[k + "_routed"]], netlist=[int(v) for v in z[k + "_netlist"]],
                   netset=[int(v) for v in z[k + "_netset"]], obs=z[k + "_obs"])


def brute_force_dist(geom, cflag, sources):
    """Independent restatement of the SPEC cost model as a dense
    numpy Bellman-Ford over all six moves; cflag uint8 [Z,Y,X] bit0 rs, bit1 fs, bit2 blk."""
    Z, Y, X = cflag.shape
    INF = np.int64(1) << 40
    rs, fs, blk = (cflag & 1).astype(np.int64), ((cflag >> 1) & 1).astype(np.int64), ((cflag >> 2) & 1).astype(np.int64)
    base = 1 + geom.drc_cost * rs + geom.fixed_shape_cost * fs
    pen = blk * (geom.block_cost * geom.layer_min_width.astype(np.int64)[:, None, None] * 20)
    horiz = (geom.layer_dir == 0).astype(np.int64)[:, None, None]
    multx = base + geom.grid_cost * (1 - horiz)
    multy = base + geom.grid_cost * horiz
    dx = np.diff(geom.x_coords.astype(np.int64))
    dy = np.diff(geom.y_coords.astype(np.int64))
    vlen = geom.via_cost * geom.layer_pitch.astype(np.int64)[1:]
    d = np.full((Z, Y, X), INF, np.int64)
    for (x, y, z) in sources:
        d[z, y, x] = 0
    while True:
        nd = d.copy()
        # +x: entering x from x-1 ; -x: entering x from x+1
        nd[:, :, 1:] = np.minimum(nd[:, :, 1:], d[:, :, :-1] + dx[None, None, :] * multx[:, :, 1:] + pen[:, :, 1:])
        nd[:, :, :-1] = np.minimum(nd[:, :, :-1], d[:, :, 1:] + dx[None, None, :] * multx[:, :, :-1] + pen[:, :, :-1])
        nd[:, 1:, :] = np.minimum(nd[:, 1:, :], d[:, :-1, :] + dy[None, :, None] * multy[:, 1:, :] + pen[:, 1:, :])
        nd[:, :-1, :] = np.minimum(nd[:, :-1, :], d[:, 1:, :] + dy[None, :, None] * multy[:, :-1, :] + pen[:, :-1, :])
        nd[1:] = np.minimum(nd[1:], d[:-1] + vlen[:, None, None] * base[1:] + pen[1:])
        nd[:-1] = np.minimum(nd[:-1], d[1:] + vlen[:, None, None] * base[:-1] + pen[:-1])
        if np.array_equal(nd, d):
            return d
        d = nd


def spec_route_net(geom, inst, usage, owner, net, goal_directed=False):
    """Second, independent restatement of SPEC steps 1-3 in plain Python on top of the dense
    Bellman-Ford above -- no code shared with oracle/xr_oracle.c: source pin, all-targets search, canonical target,
    canonical backtrace on the converged field, commit.  usage / owner [Z,Y,X] are updated in place.
    Returns (cells, conn_off, conn_cost, d_wirelength, d_via) like OracleEnv.last_paths().

    goal_directed=True replaces the converged field by a PARTIAL one: best-first search on f = d + h, h = L1 track
    distance (DBU) to the nearest unconnected access point, stopped once the smallest open f exceeds the best target
    distance B.  Only cells with d + h <= B hold their final distance; the others keep whatever tentative value (or
    infinity) the search left.  DESIGN.md section 12 argues that target choice and canonical walk come out the same
    (h is consistent: cost >= 1 per DBU, vias cost > 0; every cell of a shortest path to a best target, and every
    predecessor the walk may accept, has d + h <= B; a tentative value is never smaller than the final one, so it can
    not fake the equality d[p] + w = d[c]); the tests hold it to that."""
    Z, Y, X = usage.shape
    xc, yc = geom.x_coords.astype(np.int64), geom.y_coords.astype(np.int64)
    aps = [(int(p), tuple(int(v) for v in xyz)) for n, p, xyz in zip(inst.ap_net, inst.ap_pin, inst.ap_xyz) if n == net]
    pins = sorted(set(p for p, _ in aps))
    if len(pins) < 2:
        return [], [0], [], 0, 0
    apnet = np.zeros((Z, Y, X), np.int64)
    for n, (x, y, z) in zip(inst.ap_net, inst.ap_xyz):
        apnet[z, y, x] = n
    blk = np.zeros((Z, Y, X), np.uint8)
    for x, y, z in inst.block_xyz:
        blk[z, y, x] = 1
    cflag = ((usage > 0).astype(np.uint8) | (((apnet != 0) & (apnet != net)).astype(np.uint8) << 1) | (blk << 2))

    def w(p, c):                                          # weight of entering c from p
        (px, py, pz), (cx, cy, cz) = p, c
        f = int(cflag[cz, cy, cx])
        mult = 1 + geom.drc_cost * (f & 1) + geom.fixed_shape_cost * ((f >> 1) & 1)
        pen = geom.block_cost * int(geom.layer_min_width[cz]) * 20 * ((f >> 2) & 1)
        if pz != cz:
            return geom.via_cost * int(geom.layer_pitch[max(pz, cz)]) * mult + pen
        axis = 0 if px != cx else 1
        length = abs(int(xc[cx] - xc[px])) + abs(int(yc[cy] - yc[py]))
        return length * (mult + geom.grid_cost * (axis != int(geom.layer_dir[cz]))) + pen

    DELTA = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    xs, ys = [a[1][0] for a in aps], [a[1][1] for a in aps]
    cx2, cy2 = min(xs) + max(xs), min(ys) + max(ys)
    src_pin = min((abs(2 * x - cx2) + abs(2 * y - cy2), p) for p, (x, y, z) in aps)[1]
    tree = [xyz for p, xyz in aps if p == src_pin]
    connected = {src_pin}
    cells, off, costs, wl, via, first = [], [0], [], 0, 0, True
    def partial_field(tree, targets):
        import heapq
        INF = np.int64(1) << 40
        d = np.full((Z, Y, X), INF, np.int64)
        txy = np.array([(int(xc[x]), int(yc[y])) for (x, y, z) in targets], np.int64)
        hval = lambda c: int((np.abs(txy[:, 0] - int(xc[c[0]])) + np.abs(txy[:, 1] - int(yc[c[1]]))).min())
        tset, best, heap = set(targets), INF, []
        for c in tree:
            d[c[2], c[1], c[0]] = 0
            heapq.heappush(heap, (hval(c), 0, c))
            if c in tset:
                best = 0
        while heap:
            f, dc, c = heapq.heappop(heap)
            if f > best:
                break
            if dc != d[c[2], c[1], c[0]]:
                continue
            for k in range(6):
                v = (c[0] + DELTA[k][0], c[1] + DELTA[k][1], c[2] + DELTA[k][2])
                if not (0 <= v[0] < X and 0 <= v[1] < Y and 0 <= v[2] < Z):
                    continue
                nd = dc + w(c, v)
                if nd < d[v[2], v[1], v[0]]:
                    d[v[2], v[1], v[0]] = nd
                    heapq.heappush(heap, (nd + hval(v), nd, v))
                    if v in tset and nd < best:
                        best = nd
        return d

    while len(connected) < len(pins):
        if goal_directed:
            d = partial_field(tree, [xyz for p, xyz in aps if p not in connected])
        else:
            d = brute_force_dist(geom, cflag, tree)
        tgt = min((int(d[z, y, x]), (z * Y + y) * X + x, (x, y, z)) for p, (x, y, z) in aps if p not in connected)
        costs.append(tgt[0])
        c, last, path = tgt[2], None, [tgt[2]]
        while d[c[2], c[1], c[0]] != 0:
            for k in ([last] if last is not None else []) + list(range(6)):
                p = (c[0] - DELTA[k][0], c[1] - DELTA[k][1], c[2] - DELTA[k][2])
                if 0 <= p[0] < X and 0 <= p[1] < Y and 0 <= p[2] < Z and d[p[2], p[1], p[0]] + w(p, c) == d[c[2], c[1], c[0]]:
                    break
            else:
                raise AssertionError("no predecessor")
            c, last = p, k
            path.append(c)
        for u, v in zip(path[:-1], path[1:]):
            via += u[2] != v[2]
            wl += abs(int(xc[u[0]] - xc[v[0]])) + abs(int(yc[u[1]] - yc[v[1]]))
        for (x, y, z) in (path if first else path[:-1]):
            usage[z, y, x] = min(255, int(usage[z, y, x]) + 1)
            if owner[z, y, x] == 0:
                owner[z, y, x] = net
        tree = (path if first else tree + path)
        first = False
        on = set(tree)
        connected |= {p for p, xyz in aps if xyz in on}
        cells += [(z * Y + y) * X + x for (x, y, z) in path]
        off.append(len(cells))
    return cells, off, costs, wl, via


class OracleBackend:
    """Simulator state for xroute_env_b200.wire.SimulatorServer backed by the CPU oracle
    (test infrastructure: lets the wire layer be exercised without a GPU)."""

    def __init__(self, geom, inst):
        from oracle.oracle import OracleEnv
        self.geom, self.inst = geom, inst
        self.env = OracleEnv(geom, inst)
        self.cum = [0, 0, 0]

    def reset(self):
        self.env.reset()
        self.cum = [0, 0, 0]

    def remaining(self):
        return self.env.remaining()

    def step(self, net_id):
        m = self.env.step(int(net_id))
        self.cum = [m["violation"], m["wirelength"], m["via"]]

    def usage(self):
        return self.env.state()[0]


class WireClient:
    """Agent side of the protocol with this repo's codec (what reference Game.reset/step do on
    the wire, baseline_utils.py:392-481): REP socket for the data, REQ for the 'initial' request."""

    def __init__(self, data_port, ctrl_port):
        import zmq
        self.zmq = zmq
        self.ctx = zmq.Context()
        self.rep = self.ctx.socket(zmq.REP)
        self.rep.bind(f"tcp://127.0.0.1:{data_port}")
        self.ctrl_port = ctrl_port

    def reset(self):
        from xroute_env_b200.wire import decode_message
        req = self.ctx.socket(self.zmq.REQ)
        req.setsockopt(self.zmq.LINGER, 0)
        req.connect(f"tcp://127.0.0.1:{self.ctrl_port}")
        req.send(b"initial")
        kind, msg = decode_message(self.rep.recv())
        req.close(0)
        assert kind == "request"
        return msg

    def step(self, net_id):
        from xroute_env_b200.wire import decode_message, encode_response
        self.rep.send(encode_response(int(net_id) - 1))
        kind, msg = decode_message(self.rep.recv())
        assert kind == "request"
        if msg["is_done"]:
            self.rep.send(b"\0")
        return msg

    def close(self):
        self.rep.close(0)
        self.ctx.term()
