"""Row f2: the simulator side of the reference's ZMQ + protobuf protocol.  Codec against known
bytes and (when the reference is mounted) against its generated protobuf module; the server
against this repo's protocol client and against the UNMODIFIED reference Game."""
import contextlib
import io
import os
import socket
import sys
import threading

import numpy as np
import pytest

from helpers import OracleBackend, WireClient
from xroute_env_b200.instances import ispd18_geometry, make_instance
from xroute_env_b200.wire import (ACCESS, BLOCKAGE, NORMAL, BatchDispatcher, SimulatorServer, decode_message,
                                  encode_request, encode_response, region_nodes, request_to_data)

REF = "/root/reference/baseline"
zmq = pytest.importorskip("zmq")


def free_ports(n):
    socks = [socket.socket() for _ in range(n)]
    for s in socks:
        s.bind(("127.0.0.1", 0))
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def random_request(rng, n):
    nodes = {"maze": rng.integers(-3, 300, (n, 3)), "point": rng.integers(-5, 400000, (n, 3)),
             "type": rng.integers(0, 3, n), "is_used": rng.integers(0, 2, n),
             "net": rng.integers(-1, 70, n), "pin": rng.integers(-1, 40, n)}
    return dict(dims=[int(v) for v in rng.integers(0, 300, 3)], nodes=nodes,
                metrics=[int(v) for v in rng.integers(0, 2 ** 31, 3)], nets=[int(v) for v in rng.integers(0, 500, 7)],
                is_done=bool(rng.integers(0, 2)))


def test_codec_known_bytes():
    # Response{net_index = 2}: field 1 sint32 zigzag(2) = 4 -> 08 04, wrapped as Message field 2
    assert encode_response(2) == bytes([0x12, 0x02, 0x08, 0x04])
    assert encode_response(-1) == bytes([0x12, 0x02, 0x08, 0x01])
    assert encode_response(0) == bytes([0x12, 0x00])                       # proto3 omits the default
    assert decode_message(bytes([0x12, 0x02, 0x08, 0x01])) == ("response", -1)
    assert decode_message(bytes([0x12, 0x00])) == ("response", 0)
    # Request{dim_x=3, nodes=[Node{maze_x=1, type=ACCESS, is_used, net=0 (omitted), pin=-1}], nets=[0, 300]}
    raw = encode_request((3, 0, 0), {"maze": [[1, 0, 0]], "point": [[0, 0, 0]], "type": [ACCESS], "is_used": [1],
                                     "net": [0], "pin": [-1]}, (0, 0, 0), [0, 300])
    assert raw == bytes([0x0A, 0x11, 0x08, 0x03, 0x22, 0x08, 0x08, 0x02, 0x38, 0x02, 0x40, 0x01, 0x50, 0x01,
                         0x4A, 0x03, 0x00, 0xAC, 0x02])


def test_codec_round_trip():
    rng = np.random.default_rng(5)
    for n in (0, 1, 257):
        req = random_request(rng, n)
        kind, got = decode_message(encode_request(req["dims"], req["nodes"], req["metrics"], req["nets"], req["is_done"]))
        assert kind == "request" and got["dims"] == req["dims"] and got["metrics"] == req["metrics"]
        assert got["nets"] == req["nets"] and got["is_done"] == req["is_done"]
        for k in ("maze", "point", "type", "is_used", "net", "pin"):
            assert np.array_equal(got["nodes"][k], np.asarray(req["nodes"][k], np.int64).reshape(got["nodes"][k].shape)), k


@pytest.mark.skipif(not os.path.exists(REF), reason="reference not mounted")
def test_codec_against_reference_protobuf_module():
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    sys.path.insert(0, REF)
    import openroad_api.proto.net_ordering_pb2 as pb
    rng = np.random.default_rng(6)
    req = random_request(rng, 100)
    m = pb.Message()
    m.ParseFromString(encode_request(req["dims"], req["nodes"], req["metrics"], req["nets"], req["is_done"]))
    r = m.request
    assert [r.dim_x, r.dim_y, r.dim_z] == req["dims"] and list(r.nets) == req["nets"] and r.is_done == req["is_done"]
    assert [r.reward_violation, r.reward_wire_length, r.reward_via] == req["metrics"]
    assert len(r.nodes) == 100
    for i, nd in enumerate(r.nodes):
        assert [nd.maze_x, nd.maze_y, nd.maze_z] == list(req["nodes"]["maze"][i])
        assert [nd.point_x, nd.point_y, nd.point_z] == list(req["nodes"]["point"][i])
        assert (nd.type, int(nd.is_used), nd.net, nd.pin) == tuple(int(req["nodes"][k][i]) for k in ("type", "is_used", "net", "pin"))
    # and the other way: bytes of the reference module through this decoder
    kind, got = decode_message(m.SerializeToString())
    assert kind == "request" and got["nets"] == req["nets"] and np.array_equal(got["nodes"]["net"], req["nodes"]["net"])
    m2 = pb.Message(); m2.response.net_index = -1
    assert decode_message(m2.SerializeToString()) == ("response", -1) and m2.SerializeToString() == encode_response(-1)
    m3 = pb.Message(); m3.ParseFromString(encode_response(41))
    assert m3.HasField("response") and m3.response.net_index == 41
    # request_to_data == the reference's own decode of the same bytes
    import baseline_utils as bu

    class _Sock:
        sent = []
        def send(self, b): self.sent.append(b)
    with contextlib.redirect_stdout(io.StringIO()):
        data = bu.handle_messange(m, _Sock())
    assert data == request_to_data(got)
    assert _Sock.sent == ([b"\0"] if req["is_done"] else [])


def test_region_nodes_dense_and_sparse():
    g = ispd18_geometry(7, 6, 3)
    inst = make_instance(g, 3, 11, p_obstacle=0.2)
    usage = np.zeros((g.Z, g.Y, g.X), np.uint8)
    taken = {tuple(c) for c in inst.block_xyz.tolist()} | {tuple(c) for c in inst.ap_xyz.tolist()}
    x, y, z = next((x, y, z) for z in range(g.Z) for y in range(g.Y) for x in range(g.X) if (x, y, z) not in taken)
    usage[z, y, x] = 1
    dense, sparse = region_nodes(g, inst, usage, True), region_nodes(g, inst, usage, False)
    assert len(dense["type"]) == g.cells and (dense["type"] == BLOCKAGE).sum() == len(inst.block_xyz)
    assert (dense["type"] == ACCESS).sum() == len(inst.ap_net) and len(sparse["type"]) < g.cells
    assert np.all(dense["is_used"][dense["type"] == BLOCKAGE] == 1) and dense["is_used"].sum() == len(inst.block_xyz) + 1
    acc = dense["type"] == ACCESS
    assert dense["net"][acc].min() == 0 and np.all(dense["net"][~acc] == -1) and np.all(dense["pin"][~acc] == -1)
    assert np.array_equal(dense["point"][:, 0], g.x_coords[dense["maze"][:, 0]])
    keep = (dense["type"] != NORMAL) | (dense["is_used"] == 1)
    for k in dense:
        assert np.array_equal(dense[k][keep], sparse[k]), k


def test_server_episode_with_protocol_client():
    from oracle.oracle import OracleEnv
    g = ispd18_geometry(14, 12, 4)
    inst = make_instance(g, 5, 21)
    dp, cp = free_ports(2)
    srv = SimulatorServer(OracleBackend(g, inst), data_port=dp, ctrl_port=cp, dense=False).start()
    cli = WireClient(dp, cp)
    try:
        for episode in range(2):                          # the second 'initial' starts a fresh episode
            orc = OracleEnv(g, inst)
            msg = cli.reset()
            assert msg["dims"] == [g.X, g.Y, g.Z] and msg["metrics"] == [0, 0, 0] and not msg["is_done"]
            assert [n + 1 for n in msg["nets"]] == inst.net_ids
            order = list(np.random.default_rng(episode).permutation(inst.net_ids))
            for k, net in enumerate(order):
                msg = cli.step(net)
                m = orc.step(int(net))
                assert msg["metrics"] == [m["violation"], m["wirelength"], m["via"]]
                assert [n + 1 for n in msg["nets"]] == orc.remaining() and msg["is_done"] == (k == len(order) - 1)
                used = {tuple(c) for c, u, t in zip(msg["nodes"]["maze"].tolist(), msg["nodes"]["is_used"], msg["nodes"]["type"])
                        if u and t != BLOCKAGE}
                uz, uy, ux = np.nonzero(orc.state()[0])
                assert used == set(zip(ux.tolist(), uy.tolist(), uz.tolist()))
        assert srv.episodes == 2 and srv.steps == 2 * len(inst.net_ids)
    finally:
        cli.close(); srv.close()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference not mounted")
def test_unmodified_reference_game_runs_against_the_server():
    """The reference's own Game (ZMQ REP + protobuf + build_3Dgrid) plays whole episodes against
    SimulatorServer; what it returns equals the oracle stepped directly, observation included."""
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    sys.path.insert(0, REF)
    import baseline_utils as bu
    from oracle.oracle import OracleEnv
    g = ispd18_geometry(10, 9, 5)
    inst = make_instance(g, 4, 33)
    dp, cp = free_ports(2)
    srv = SimulatorServer(OracleBackend(g, inst), data_port=dp, ctrl_port=cp).start()
    game = bu.Game(port_recv=str(dp), port_initial=str(cp))
    try:
        for episode in range(2):
            orc = OracleEnv(g, inst)
            with contextlib.redirect_stdout(io.StringIO()):
                obs, tries = game.reset()
            assert tries == 0 and set(game.action_space) == set(inst.net_ids)
            assert np.array_equal(obs.numpy(), orc.obs())
            for net in np.random.default_rng(10 + episode).permutation(inst.net_ids):
                with contextlib.redirect_stdout(io.StringIO()):
                    obs, done, vio, wl, via = game.step(int(net))
                m = orc.step(int(net))
                assert (vio, wl, via, done) == (m["d_violation"], m["d_wirelength"], m["d_via"], bool(m["done"]))
                assert np.array_equal(obs.numpy(), orc.obs()) and set(game.legal_action_set) == set(orc.remaining())
            assert done
    finally:
        game.socket.close(0); srv.close()


def test_batch_dispatcher_gathers_concurrent_actions():
    class FakeVg:
        n_envs = 4
        def __init__(self): self.calls = []
        def step(self, acts): self.calls.append(acts.copy())
        def results_host(self):
            return None, None, np.cumsum(np.stack(self.calls), 0)[-1][:, None].repeat(6, 1)
    vg = FakeVg()
    d = BatchDispatcher(vg, max_wait_s=0.05)
    out = {}
    ts = [threading.Thread(target=lambda e=e: out.__setitem__(e, d.submit(e, e + 1))) for e in range(4)]
    for t in ts: t.start()
    for t in ts: t.join(timeout=5)
    d.close()
    assert d.batches < 4, "concurrent submissions must share batched steps"
    assert sum(int(c.sum()) for c in vg.calls) == 1 + 2 + 3 + 4 and out == {e: [e + 1] * 3 for e in range(4)}


def test_evaluation_mode_drives_a_bare_rep_loop():
    """The protocol of the reference's inference servers (baseline/PPO/test_PPO.py:50-86): the agent only answers
    Requests on a REP socket; the simulator side starts by itself, never sends is_done and moves on to the next
    episode when no net is left."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import build_3Dgrid  # noqa: F401  (the GPU drop-in is exercised in tests/test_gpu_wire.py)
    g = ispd18_geometry(12, 11, 4)
    inst = make_instance(g, 4, 44)
    dp, = free_ports(1)
    ctx = zmq.Context()
    rep = ctx.socket(zmq.REP)
    rep.bind(f"tcp://127.0.0.1:{dp}")
    srv = SimulatorServer(OracleBackend(g, inst), data_port=dp, evaluation=True, max_episodes=2, dense=False).start()
    try:
        poller = zmq.Poller(); poller.register(rep, zmq.POLLIN)
        seen = []
        orc = OracleEnv(g, inst)
        for k in range(2 * len(inst.net_ids)):
            assert poller.poll(5000), "no request from the simulator side"
            kind, msg = decode_message(rep.recv())
            assert kind == "request" and not msg["is_done"]
            if k % len(inst.net_ids) == 0:
                orc.reset()
                assert msg["metrics"] == [0, 0, 0]
            assert [n + 1 for n in msg["nets"]] == orc.remaining()
            data = request_to_data(msg)                                  # what handle_messange would hand to build_3Dgrid
            assert data[3] == orc.remaining() and data[0] == [g.X, g.Y, g.Z]
            net = msg["nets"][-1] + 1                                    # "policy": always the largest remaining id
            seen.append(net)
            orc.step(net)
            rep.send(encode_response(net - 1))
        assert not poller.poll(300)                                      # two episodes played, the server is done
        assert srv.episodes == 2 and srv.steps == 2 * len(inst.net_ids) and seen[: len(inst.net_ids)] == sorted(inst.net_ids, reverse=True)
    finally:
        srv.close(); rep.close(0); ctx.term()
