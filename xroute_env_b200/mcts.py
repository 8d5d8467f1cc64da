"""Graph observation and prefix-re-route step semantics of the reference's MCTS flavour.

The MuZero-style baseline (``/root/reference/baseline/xroute``) does not route one net per step.  Its dispatcher
(``trainer4/dispatcher.py:49-122``) keeps a list of nets the agent has chosen so far; every step it RE-ROUTES THE WHOLE
REGION in the order ``chosen + remaining (default order)``, reads the cumulative metrics of that complete order, turns
them into deltas against the previous step (``:73-81``: PREVIOUS minus current, so an order that got cheaper yields
positive numbers; the very first observation carries minus the cost of the default order), marks the chosen nets in the graph observation (node property 3 = ``is_routed``, ``:84-85``) and hands the agent
the nets still to choose.  The agent's reward is ``(0.5 wl + 4 via + 500 vio) / 1000`` of the deltas
(``net_order.py:198``); the observation is a graph -- one node per net with 11 features, one edge per pair of nets that
overlap (``net_ordering.proto:30-41,66-80``; consumed as ``x`` / ``edge_index`` of a GCN, ``self_route.py:300-301``).

``PrefixRerouteGame`` restates those step semantics, batched over the environments of a ``VecGame`` (every step is
``reset`` + one batched route of the whole order).  Its bookkeeping is pinned by ``tests/golden/mcts_dispatch.npz``: the
messages the reference's own ``Dispatcher.run`` produces when it is driven with a fake mixer.

Node features.  The reference computes them inside the absent simulator binary; the proto comment only names three
families ("pin_nums, access_point_ratios, region_volume_ratios") and the dispatcher fixes slot 3.  The layout below is
THIS repository's definition (unpinned), 11 floats per net, 0-based net index = net id - 1:
    [0] pins   [1] access points per pin   [2] the net's share of the region's access points   [3] is_routed
    [4] box width / X   [5] box height / Y   [6] box layers / Z   [7] box area / (X Y)   [8] box volume / (X Y Z)
    [9] box centre x / X   [10] box centre y / Y            (box = bounding box of the net's access points, in cells)
Edges: every unordered pair of nets whose boxes overlap in x, y and z, as ``[i, j]`` with ``i < j`` (0-based).
"""
from __future__ import annotations

import numpy as np

N_FEATURES = 11


def graph_features(geom, inst):
    """(node float32 [n, 11] for net ids 1..n with n = largest net id, edges int32 [E, 2]).  Rows of ids without access
    points are zero; slot 3 (is_routed) is left 0."""
    n = int(inst.ap_net.max()) if len(inst.ap_net) else 0
    node = np.zeros((n, N_FEATURES), np.float32)
    box = np.zeros((n, 6), np.int64)
    has = np.zeros(n, bool)
    total_ap = max(len(inst.ap_net), 1)
    X, Y, Z = geom.X, geom.Y, geom.Z
    for k in range(1, n + 1):
        m = inst.ap_net == k
        if not m.any():
            continue
        xyz = inst.ap_xyz[m]
        lo, hi = xyz.min(0), xyz.max(0)
        w, h, l = (hi - lo + 1).tolist()
        pins = len(set(inst.ap_pin[m].tolist()))
        has[k - 1] = True
        box[k - 1] = [lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]]
        node[k - 1] = [pins, m.sum() / pins, m.sum() / total_ap, 0.0, w / X, h / Y, l / Z, w * h / (X * Y),
                       w * h * l / (X * Y * Z), (lo[0] + hi[0] + 1) / 2 / X, (lo[1] + hi[1] + 1) / 2 / Y]
    edges = []
    for i in range(n):
        if not has[i]:
            continue
        for j in range(i + 1, n):
            if has[j] and box[i, 0] <= box[j, 1] and box[j, 0] <= box[i, 1] and box[i, 2] <= box[j, 3] and \
                    box[j, 2] <= box[i, 3] and box[i, 4] <= box[j, 5] and box[j, 4] <= box[i, 5]:
                edges.append((i, j))
    return node, np.asarray(edges, np.int32).reshape(-1, 2)


def mcts_reward(d_violation, d_wirelength, d_via):
    """``net_order.py:198``."""
    return (0.5 * d_wirelength + 4 * d_via + 500 * d_violation) / 1000


class PrefixRerouteGame:
    """Step semantics of ``trainer4/dispatcher.py:49-122`` for a batch of regions.

    ``route_order(orders int32 [n_steps, N]) -> cumulative metrics [N, >=3]`` (violation, wirelength, via) must route
    every environment from scratch in the given order (0 = idle): ``VecGame.route_order`` on the GPU, or any stand-in.
    ``net_ids[e]``: the net ids (1-based, ascending = the simulator's default order) of environment ``e``.
    Actions and ``unrouted`` lists use the reference's 0-based net indices (``Response.net_index``)."""

    def __init__(self, route_order, net_ids, init_metrics=None):
        self._route = route_order
        self.N = len(net_ids)
        self.default = [[int(k) - 1 for k in ids] for ids in net_ids]
        self.init = np.zeros((self.N, 3), np.int64) if init_metrics is None else np.asarray(init_metrics, np.int64)
        self.reset()

    def _cost(self):
        n = max((len(r) + len(u) for r, u in zip(self.routed, self.unrouted)), default=0)
        orders = np.zeros((max(n, 1), self.N), np.int32)
        for e in range(self.N):
            seq = self.routed[e] + self.unrouted[e]
            orders[:len(seq), e] = np.asarray(seq, np.int32) + 1
        cum = self._route(orders)
        cum = cum.cpu().numpy() if hasattr(cum, "cpu") else np.asarray(cum)
        return cum[:, :3].astype(np.int64)

    def reset(self):
        """First observation: the deltas are the cost of the default order (``dispatcher.py:53-66,73-81``)."""
        self.routed = [[] for _ in range(self.N)]
        self.unrouted = [list(d) for d in self.default]
        self.last = np.zeros((self.N, 3), np.int64)
        return self._observe()

    def _observe(self):
        current = self._cost() - self.init
        delta = self.last - current          # dispatcher.py:76: [b - a for a, b in zip(current, last)] = previous minus current
        self.last = current
        done = np.array([len(u) == 0 for u in self.unrouted])
        return {"delta": delta, "done": done, "nets": [list(u) for u in self.unrouted],
                "is_routed": [sorted(r) for r in self.routed],
                "reward": mcts_reward(delta[:, 0], delta[:, 1], delta[:, 2])}

    def step(self, actions):
        """``actions[e]``: 0-based index of the net environment ``e`` routes next (must be in ``nets[e]``), or -1 to
        leave a finished environment alone."""
        for e, a in enumerate(actions):
            a = int(a)
            if a < 0:
                continue
            if a not in self.unrouted[e]:
                raise ValueError(f"environment {e}: net {a} is not in the unrouted set")
            self.routed[e].append(a)
            self.unrouted[e].remove(a)
        return self._observe()


def graph_observation(vec_game):
    """Graph observation of every environment of a ``VecGame`` as device tensors: ``(x float32 [N, max_nets, 11] with slot
    3 = is_routed read from the library's ``legal`` buffer, [edge_index int64 [2, E_e]] per environment)``."""
    import torch
    nodes = np.zeros((vec_game.n_envs, vec_game.max_nets, N_FEATURES), np.float32)
    edges, present = [], np.zeros((vec_game.n_envs, vec_game.max_nets), bool)
    for e, inst in enumerate(vec_game.insts):
        x, ed = graph_features(vec_game.geom, inst)
        nodes[e, :len(x)] = x
        present[e, :len(x)] = x[:, 0] > 0
        edges.append(torch.from_numpy(ed.astype(np.int64).T.copy()).cuda(vec_game.device))
    x = torch.from_numpy(nodes).cuda(vec_game.device)
    legal = vec_game.legal[:, 1:].bool()                                       # net ids 1..max_nets still to route
    x[:, :, 3] = (torch.from_numpy(present).cuda(vec_game.device) & ~legal).float()
    return x, edges
