"""Row f3, MCTS flavour: prefix-re-route step semantics and the graph observation.

tests/golden/mcts_dispatch.npz holds the messages the reference's own dispatcher (baseline/xroute/trainer4/dispatcher.py,
unmodified, run by tests/golden/make_golden.py with a fake mixer on the CPU oracle) hands its algorithm; PrefixRerouteGame
must reproduce them -- on the CPU with the oracle as router, on the GPU with VecGame.route_order."""
import os

import numpy as np
import pytest

from helpers import GOLD
from xroute_env_b200.instances import Instance, ispd18_geometry, make_instance
from xroute_env_b200.mcts import N_FEATURES, PrefixRerouteGame, graph_features, mcts_reward


def _golden():
    z = np.load(os.path.join(GOLD, "mcts_dispatch.npz"))
    geom = ispd18_geometry(*[int(v) for v in z["dims"]])
    inst = Instance(block_xyz=z["block_xyz"], ap_net=z["ap_net"], ap_pin=z["ap_pin"], ap_xyz=z["ap_xyz"])
    return z, geom, inst


def _check(game, z):
    obs = game.reset()
    for k in range(len(z["delta"])):
        assert obs["delta"][0].tolist() == z["delta"][k].tolist(), k
        assert obs["nets"][0] == [int(v) for v in z["nets"][k] if v >= 0], k
        flags = [1 if i in obs["is_routed"][0] else 0 for i in range(z["is_routed"].shape[1])]
        assert flags == z["is_routed"][k].tolist(), k
        assert bool(obs["done"][0]) == bool(z["is_done"][k]), k
        d = z["delta"][k]
        assert obs["reward"][0] == (0.5 * d[1] + 4 * d[2] + 500 * d[0]) / 1000 == mcts_reward(d[0], d[1], d[2])
        if k < len(z["choices"]):
            obs = game.step([int(z["choices"][k])])


def test_prefix_reroute_matches_the_reference_dispatcher_cpu():
    from oracle.oracle import OracleEnv
    z, geom, inst = _golden()
    env = OracleEnv(geom, inst)

    def route_order(orders):
        env.reset()
        m = {"violation": 0, "wirelength": 0, "via": 0}
        for net in orders[:, 0]:
            if net:
                m = env.step(int(net))
        return np.array([[m["violation"], m["wirelength"], m["via"]]])
    _check(PrefixRerouteGame(route_order, [inst.net_ids]), z)


def test_graph_features_layout():
    geom = ispd18_geometry(20, 18, 4)
    inst = Instance(block_xyz=np.zeros((0, 3), np.int32), ap_net=np.array([1, 1, 1, 2, 2, 4, 4], np.int32),
                    ap_pin=np.array([1, 1, 2, 1, 2, 1, 2], np.int32),
                    ap_xyz=np.array([[2, 3, 0], [3, 3, 0], [9, 7, 1], [5, 5, 0], [12, 6, 0], [15, 15, 1], [18, 16, 1]], np.int32))
    x, e = graph_features(geom, inst)
    assert x.shape == (4, N_FEATURES) and not x[2].any()                     # net 3 has no access points
    assert np.allclose(x[0, :4], [2.0, 1.5, 3 / 7, 0.0])
    assert np.allclose(x[0, 4:9], [8 / 20, 5 / 18, 2 / 4, 40 / 360, 80 / 1440])
    assert np.allclose(x[0, 9:], [6.0 / 20, 5.5 / 18])
    assert e.tolist() == [[0, 1]]                                            # boxes of nets 1 and 2 overlap; net 4 is apart


@pytest.mark.gpu
def test_prefix_reroute_and_graph_observation_on_the_gpu():
    import torch
    from xroute_env_b200 import VecGame
    from xroute_env_b200.mcts import graph_observation
    z, geom, inst = _golden()
    other = make_instance(geom, 5, 99)
    vg = VecGame(geom, [inst, other], device=0)
    game = PrefixRerouteGame(vg.route_order, [inst.net_ids, other.net_ids])
    obs = game.reset()
    for k in range(len(z["delta"])):
        assert obs["delta"][0].tolist() == z["delta"][k].tolist(), k
        assert obs["nets"][0] == [int(v) for v in z["nets"][k] if v >= 0]
        if k < len(z["choices"]):
            second = obs["nets"][1][0] if obs["nets"][1] else -1
            obs = game.step([int(z["choices"][k]), second])
    assert obs["done"].all()
    # graph observation from the device buffers: after routing nets 3 and 1 of environment 0 incrementally
    vg.reset()
    vg.step(np.array([3, 0], np.int32)); vg.step(np.array([1, 2], np.int32))
    x, edges = graph_observation(vg)
    want, e0 = graph_features(geom, inst)
    assert x.is_cuda and x.shape == (2, vg.max_nets, N_FEATURES)
    assert x[0, :, 3].cpu().tolist()[:7] == [1, 0, 1, 0, 0, 0, 0] and x[1, :, 3].cpu().tolist()[:5] == [0, 1, 0, 0, 0]
    xs = x[0, :len(want)].cpu().numpy().copy(); xs[:, 3] = 0
    assert np.array_equal(xs, want) and np.array_equal(edges[0].cpu().numpy().T, e0)
    vg.close()
