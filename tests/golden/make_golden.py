#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING the reference.

Needs /root/reference (read-only mount of xrouting/xroute_env); it is only ever run in
the build container, never on the GPU box.  Nothing from the reference is copied: the
fixtures hold inputs (seeded node lists) and the outputs the reference computes.

  obs_cases.npz        observations of the reference's own build_3Dgrid
                       (baseline/build_3Dgrid.py:224-270) on seeded node lists, in
                       training mode (several routed sets) and inference mode.
  reward_tfevents.npz  the 717 per-episode (violation, wirelength, via, reward) tuples
                       of the shipped PPO log
                       (baseline/PPO/results/2023-04-27--05-00-38/events.out.tfevents.*),
                       the known-answer vectors of train_PPO.py:101-102.
  mcts_dispatch.npz    the messages of the reference's own MCTS dispatcher (trainer4/dispatcher.py)
                       driven with a fake mixer: prefix-re-route bookkeeping, delta metrics, is_routed.
  game_episode.npz     a 2-net episode driven through the UNMODIFIED reference Game
                       (baseline/baseline_utils.py:383-481) against a fake REQ-side
                       simulator: pins the cumulative->delta differencing, done flag,
                       1-based action shift and legal sets.

    python tests/golden/make_golden.py
"""
import contextlib
import glob
import io
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "baseline"))


def random_nodes(rng, X, Y, Z, n_nets, p_block, p_used, sparse):
    """A seeded node stream in the decoded layout of handle_messange
    (baseline_utils.py:23-40): rows (x, y, z, used, Net, Pin)."""
    net = np.zeros((X, Y, Z), np.int64)
    pin = np.full((X, Y, Z), -1, np.int64)
    net[rng.random((X, Y, Z)) < p_block] = -1
    for n in range(1, n_nets + 1):
        if n == 3 and n_nets >= 4:
            continue                                  # a net id with no access point at all
        for p in range(1, int(rng.integers(2, 5)) + 1):
            x, y, z = int(rng.integers(0, X)), int(rng.integers(0, Y)), int(rng.integers(0, min(2, Z)))
            for k in range(int(rng.integers(1, 4))):
                # runs along x, y or z so that every adjacency direction occurs
                ax = int(rng.integers(0, 3))
                xx, yy, zz = x + (k if ax == 0 else 0), y + (k if ax == 1 else 0), z + (k if ax == 2 else 0)
                if xx < X and yy < Y and zz < Z and net[xx, yy, zz] == 0:
                    net[xx, yy, zz] = n
                    pin[xx, yy, zz] = p
    used = (rng.random((X, Y, Z)) < p_used) | (net == -1)
    rows = []
    for x in range(X):
        for y in range(Y):
            for z in range(Z):
                if sparse and net[x, y, z] == 0 and not used[x, y, z] and rng.random() < 0.7:
                    continue                          # NORMAL unused nodes may be absent
                rows.append((x, y, z, int(used[x, y, z]), int(net[x, y, z]), int(pin[x, y, z])))
    rows = np.array(rows, np.int32)
    return rows[rng.permutation(len(rows))]            # node order must not matter


def rows_to_data(rows, dims, cum, netlist):
    nodes = [[[int(r[0]), int(r[1]), int(r[2])], [0, 0, int(r[2])], [int(r[3]), int(r[4]), int(r[5])]] for r in rows]
    return [list(dims), nodes, list(cum), list(netlist)]


def make_obs_cases():
    import build_3Dgrid as ref
    rng = np.random.default_rng(20260417)
    out = {}
    cases = [  # X, Y, Z, nets, p_block, p_used, sparse
        (5, 6, 5, 3, 0.10, 0.05, False),
        (3, 2, 5, 2, 0.00, 0.00, False),
        (12, 10, 5, 6, 0.10, 0.10, True),
        (25, 26, 9, 10, 0.08, 0.05, True),
        (7, 9, 2, 5, 0.30, 0.30, False),
        (16, 5, 3, 0, 0.20, 0.10, False),            # no nets at all: obstacle + order channels only
    ]
    idx = 0
    for (X, Y, Z, nn, pb, pu, sp) in cases:
        rows = random_nodes(rng, X, Y, Z, nn, pb, pu, sp)
        nets_present = sorted(set(int(v) for v in rows[:, 4] if v >= 1))
        variants = [("train", set(), None)]
        if nets_present:
            variants.append(("train", set(nets_present[::2]), None))
            variants.append(("train", set(nets_present), None))          # everything routed
            variants.append(("infer", set(), nets_present[1:] + [99]))    # inference: filter by data[3]
        for mode, routed, netlist in variants:
            data = rows_to_data(rows, (X, Y, Z), (3, 1010, 2), netlist if netlist is not None else nets_present)
            with contextlib.redirect_stdout(io.StringIO()):
                obs, netset, v, w, a = ref.build_3Dgrid(data, routed, bool_inference=(mode == "infer"))
            assert (v, w, a) == (3, 1010, 2)
            k = f"c{idx}"
            out[k + "_dims"] = np.array([X, Y, Z], np.int32)
            out[k + "_rows"] = rows
            out[k + "_mode"] = np.array([mode == "infer"], np.int32)
            out[k + "_routed"] = np.array(sorted(routed), np.int32)
            out[k + "_netlist"] = np.array(netlist if netlist is not None else nets_present, np.int32)
            out[k + "_netset"] = np.array(sorted(netset), np.int32)
            out[k + "_obs"] = obs.numpy()
            idx += 1
    out["n_cases"] = np.array([idx], np.int32)
    np.savez_compressed(os.path.join(HERE, "obs_cases.npz"), **out)
    print("obs_cases.npz:", idx, "cases")


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]; i += 1
        r |= (c & 0x7F) << s; s += 7
        if not c & 0x80:
            return r, i


def _fields(b):
    i = 0
    while i < len(b):
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v = b[i:i + 8]; i += 8
        elif wt == 2:
            ln, i = _varint(b, i); v = b[i:i + ln]; i += ln
        elif wt == 5:
            v = b[i:i + 4]; i += 4
        else:
            raise ValueError(wt)
        yield fn, wt, v


def make_reward_vectors():
    f = glob.glob(os.path.join(REF, "baseline/PPO/results/*/events.out.tfevents*"))[0]
    data = open(f, "rb").read()
    pos, rows = 0, {}
    while pos < len(data):
        (ln,) = struct.unpack("<Q", data[pos:pos + 8]); pos += 12
        ev = data[pos:pos + ln]; pos += ln + 4
        step = None
        for fn, wt, v in _fields(ev):
            if fn == 2 and wt == 0:
                step = v
            if fn == 5 and wt == 2:
                for fn2, _, v2 in _fields(v):
                    if fn2 != 1:
                        continue
                    tag = val = None
                    for fn3, wt3, v3 in _fields(v2):
                        if fn3 == 1:
                            tag = v3.decode()
                        if fn3 == 2 and wt3 == 5:
                            val = struct.unpack("<f", v3)[0]
                    rows.setdefault(tag, {})[step] = val
    steps = sorted(rows["1.Episode/1.reward"])
    arr = np.array([[rows["1.Episode/2.violation"][s], rows["1.Episode/3.wirelength"][s],
                     rows["1.Episode/4.via"][s], rows["1.Episode/1.reward"][s]] for s in steps], np.float64)
    np.savez_compressed(os.path.join(HERE, "reward_tfevents.npz"), tuples=arr)
    ok = np.all(arr[:, 3] == -(500 * arr[:, 0] + 4 * arr[:, 2] + 0.5 * arr[:, 1]))
    print("reward_tfevents.npz:", arr.shape, "formula holds on all rows:", bool(ok))


def make_game_episode():
    """Drive the unmodified reference Game with a fake simulator (REQ side) and record what
    it returns; ports are overridden so nothing else on the host is touched."""
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    import threading
    import zmq
    import baseline_utils as bu
    import openroad_api.proto.net_ordering_pb2 as pb

    X, Y, Z = 6, 5, 5
    aps = {0: [(0, (1, 1, 0)), (1, (4, 3, 0))], 1: [(0, (0, 4, 1)), (0, (1, 4, 1)), (1, (5, 0, 1))]}
    script = [  # (nets still to route, cumulative metrics, used cells, is_done)
        ([0, 1], (0, 0, 0), set(), False),
        ([0], (1, 1010, 2), {(0, 4, 1), (1, 4, 1), (2, 4, 1), (5, 0, 1)}, False),
        ([], (1, 2010, 4), {(0, 4, 1), (1, 4, 1), (2, 4, 1), (5, 0, 1), (1, 1, 0), (2, 1, 0), (4, 3, 0)}, True),
    ]
    PORT_DATA, PORT_CTRL = "15756", "16767"
    replies = []

    def build_req(nets, cum, used, is_done):
        m = pb.Message()
        r = m.request
        r.dim_x, r.dim_y, r.dim_z = X, Y, Z
        r.reward_violation, r.reward_wire_length, r.reward_via = cum
        r.is_done = is_done
        r.nets.extend(nets)
        apmap = {xyz: (n, p) for n, lst in aps.items() for p, xyz in lst}
        for x in range(X):
            for y in range(Y):
                for z in range(Z):
                    nd = r.nodes.add()
                    nd.maze_x, nd.maze_y, nd.maze_z = x, y, z
                    nd.point_x, nd.point_y, nd.point_z = 200 + 400 * x, 190 + 380 * y, z
                    nd.is_used = (x, y, z) in used
                    if (x, y, z) in apmap:
                        nd.type = pb.NodeType.ACCESS
                        nd.net, nd.pin = apmap[(x, y, z)]
                    elif (x + 2 * y + z) % 11 == 0:
                        nd.type = pb.NodeType.BLOCKAGE
                        nd.net = nd.pin = -1
                        nd.is_used = True
                    else:
                        nd.type = pb.NodeType.NORMAL
                        nd.net = nd.pin = -1
        return m.SerializeToString()

    def ctrl_server():
        ctx = zmq.Context()
        s = ctx.socket(zmq.REP)
        s.bind("tcp://127.0.0.1:" + PORT_CTRL)
        s.recv()
        s.send(b"\0")
        s.close(0)

    def simulator():
        ctx = zmq.Context()
        s = ctx.socket(zmq.REQ)
        s.connect("tcp://127.0.0.1:" + PORT_DATA)
        for nets, cum, used, is_done in script:
            s.send(build_req(nets, cum, used, is_done))
            rep = s.recv()
            if rep == b"\0":
                replies.append(-99)
            else:
                m = pb.Message(); m.ParseFromString(rep)
                replies.append(m.response.net_index)
        s.close(0)

    threading.Thread(target=ctrl_server, daemon=True).start()
    th = threading.Thread(target=simulator, daemon=True)
    th.start()
    game = bu.Game(port_recv=PORT_DATA, port_initial=PORT_CTRL)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        obs, tries = game.reset()
        out["obs0"] = obs.numpy(); out["tries"] = np.array([tries]); out["space0"] = np.array(sorted(game.action_space))
        o1, d1, v1, w1, a1 = game.step(2)
        out["obs1"] = o1.numpy(); out["ret1"] = np.array([int(d1), v1, w1, a1]); out["legal1"] = np.array(sorted(game.legal_action_set))
        o2, d2, v2, w2, a2 = game.step(1)
        out["obs2"] = o2.numpy(); out["ret2"] = np.array([int(d2), v2, w2, a2]); out["legal2"] = np.array(sorted(game.legal_action_set), np.int64)
    th.join(timeout=5)
    out["replies"] = np.array(replies)
    out["dims"] = np.array([X, Y, Z])
    np.savez_compressed(os.path.join(HERE, "game_episode.npz"), **out)
    print("game_episode.npz: replies", replies, "ret1", out["ret1"], "ret2", out["ret2"])


def make_a3c_features():
    """22-feature per-net vectors of the A3C flavour: the reference's own ``Game.get_feature`` and
    ``Game._cal_reward`` (baseline/A3C/utils.py:195-277) run on seeded regions.  The module imports a
    generated protobuf file that the repository does not ship (compile_proto.sh output); it is
    stubbed -- get_feature never touches it."""
    import types
    for name in ("openroad_api", "openroad_api.proto", "openroad_api.proto.net_ordering_pb2"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, os.path.join(REF, "baseline", "A3C"))
    saved = sys.modules.pop("utils", None)
    import utils as a3c
    from xroute_env_b200.instances import export_data, ispd18_geometry, make_instance
    rng = np.random.default_rng(99)
    out = {}
    cases = [(12, 10, 5, 5, 3), (25, 26, 9, 12, 4), (30, 8, 9, 7, 5), (9, 9, 2, 3, 6)]
    for idx, (X, Y, Z, n_nets, seed) in enumerate(cases):
        g = ispd18_geometry(X, Y, Z)
        inst = make_instance(g, n_nets, seed)
        data = export_data(g, inst, np.zeros((Z, Y, X), np.uint8))
        some = [int(v) for v in rng.permutation(inst.net_ids)[: max(1, n_nets // 2)]]
        count_map = {str(n): int(rng.integers(1, 4)) for n in some}
        metrics_delta = {str(n): [int(rng.integers(0, 3)), int(rng.integers(0, 50000)), int(rng.integers(0, 30))] for n in some}
        game = a3c.Game.__new__(a3c.Game)
        game.observation = None
        game.accessPoints = None
        with contextlib.redirect_stdout(io.StringIO()):
            feats = game.get_feature(data + [[], [], count_map, metrics_delta])
        nets = sorted(feats)
        k = f"c{idx}"
        out[k + "_dims"] = np.array([X, Y, Z, n_nets, seed], np.int32)
        out[k + "_nets"] = np.array(nets, np.int32)
        out[k + "_feat"] = np.stack([feats[n] for n in nets]).astype(np.float64)
        out[k + "_count"] = np.array([[int(n), c] for n, c in count_map.items()], np.int64)
        out[k + "_delta"] = np.array([[int(n)] + v for n, v in metrics_delta.items()], np.int64)
    costs = np.array([[0, 0, 0], [1, 1010, 2], [3, 123456, 77], [0, 5, 0]], np.int64)
    game = a3c.Game.__new__(a3c.Game)
    out["cost_in"] = costs
    out["cost_out"] = np.array([game._cal_reward(list(c)) for c in costs], np.float64)
    out["n_cases"] = np.array([len(cases)], np.int32)
    np.savez_compressed(os.path.join(HERE, "a3c_features.npz"), **out)
    if saved is not None:
        sys.modules["utils"] = saved
    print("a3c_features.npz:", len(cases), "cases; cost", out["cost_out"].tolist())


def make_mcts_dispatch():
    """mcts_dispatch.npz: the messages the reference's own Dispatcher.run (baseline/xroute/trainer4/dispatcher.py:36-122,
    unmodified) hands its algorithm when its mixer -- the OpenROAD process that routes a complete net order -- is replaced
    by a fake one backed by this repo's CPU oracle: pins the prefix-re-route bookkeeping (order = chosen + remaining,
    cumulative -> delta metrics with the default order's cost in the first message, is_routed in node property 3,
    remaining-net lists, is_done)."""
    import threading
    import types
    import zmq
    from oracle.oracle import OracleEnv
    from xroute_env_b200.instances import ispd18_geometry, make_instance
    from xroute_env_b200.mcts import graph_features
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    sys.path.insert(0, os.path.join(REF, "baseline", "xroute"))
    import net_ordering_pb2 as pb
    proto_pkg = types.ModuleType("proto"); proto_pkg.net_ordering_pb2 = pb
    sys.modules["proto"] = proto_pkg; sys.modules["proto.net_ordering_pb2"] = pb
    geom = ispd18_geometry(30, 28, 6)
    inst = make_instance(geom, 7, 4711, p_obstacle=0.15)
    nodes, _ = graph_features(geom, inst)
    env = OracleEnv(geom, inst)

    class FakeMixer:
        """get_observation / set_net_list / get_result / ack of trainer4/mixer.py:52-69 on the oracle"""
        pending = None

        def __init__(self, **kw):
            pass

        def start(self):
            pass

        def _request(self, cum):
            r = pb.Request()
            r.dim_x, r.dim_y, r.dim_z = geom.X, geom.Y, geom.Z
            r.reward_violation, r.reward_wire_length, r.reward_via = cum
            r.nets[:] = [n - 1 for n in inst.net_ids]
            for row in nodes:
                r.graph.node_properties.add().values[:] = [float(v) for v in row]
            return r

        def get_observation(self):
            if FakeMixer.pending is not None:
                r, FakeMixer.pending = FakeMixer.pending, None
                return r
            return self._request((0, 0, 0))

        get_result = get_observation

        def set_net_list(self, net_list):
            env.reset()
            m = None
            for k in net_list:
                m = env.step(int(k) + 1)
            FakeMixer.pending = self._request((m["violation"], m["wirelength"], m["via"]))

        def ack(self):
            pass

    mixer_mod = types.ModuleType("mixer"); mixer_mod.Mixer = FakeMixer
    sys.modules["mixer"] = mixer_mod
    sys.path.insert(0, os.path.join(REF, "baseline", "xroute", "trainer4"))
    import dispatcher as ref_dispatcher
    port = 17651
    choices = [4, 0, 6, 2, 5, 1, 3]                        # 0-based net indices, in the order the "agent" picks them
    log = []

    def algorithm():
        ctx = zmq.Context()
        sock = ctx.socket(zmq.REP)
        sock.bind(f"tcp://127.0.0.1:{port}")
        k = 0
        while True:
            msg = pb.Message(); msg.ParseFromString(sock.recv())
            q = msg.request
            log.append((list(q.nets), [q.reward_violation, q.reward_wire_length, q.reward_via],
                        [int(p.values[3]) for p in q.graph.node_properties], bool(q.is_done)))
            if q.is_done:
                sock.send(b"\0")
                break
            out = pb.Message(); out.response.net_index = choices[k]; k += 1
            sock.send(out.SerializeToString())
        sock.close(0); ctx.term()

    th = threading.Thread(target=algorithm)
    th.start()
    d = ref_dispatcher.Dispatcher(openroad_executable="none", port_to_alg=port, port_from_or_mixer=0)
    with contextlib.redirect_stdout(io.StringIO()):
        d.run()
    th.join(timeout=30)
    assert len(log) == len(choices) + 1 and log[-1][3]
    np.savez_compressed(
        os.path.join(HERE, "mcts_dispatch.npz"), dims=np.array([geom.X, geom.Y, geom.Z]), seed=np.array([4711]),
        block_xyz=inst.block_xyz, ap_net=inst.ap_net, ap_pin=inst.ap_pin, ap_xyz=inst.ap_xyz, choices=np.array(choices),
        nets=np.array([l[0] + [-1] * (7 - len(l[0])) for l in log]), delta=np.array([l[1] for l in log]),
        is_routed=np.array([l[2] for l in log]), is_done=np.array([l[3] for l in log]))
    print("mcts_dispatch.npz:", len(log), "messages; first delta (cost of the default order)", log[0][1])


if __name__ == "__main__":
    which = sys.argv[1:] or ["obs", "reward", "game", "a3c", "mcts"]
    if "obs" in which:
        make_obs_cases()
    if "reward" in which:
        make_reward_vectors()
    if "game" in which:
        make_game_episode()
    if "a3c" in which:
        make_a3c_features()
    if "mcts" in which:
        make_mcts_dispatch()
