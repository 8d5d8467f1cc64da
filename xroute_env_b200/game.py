"""Drop-in for the reference's single-environment ``Game`` and ``build_3Dgrid``.

``Game`` keeps the interface of ``/root/reference/baseline/baseline_utils.py:383-481``
(same constructor arguments, ``reset() -> (obs, reset_try_time)``,
``step(action) -> (obs, done, violation, wirelength, via)``, attributes ``routed_nets``,
``action_space``, ``legal_action_set``, ``observation``) so ``train_PPO.py:29,70,99`` /
``train_DQN.py:28,80,97`` run unchanged.  Instead of the ZMQ round trip to the OpenROAD
simulator it drives a one-environment ``VecGame`` on the GPU; regions come from an
instance source (an iterable of ``Instance``) that plays the role of the launcher
cycling through ``dump/worker*`` directories (``examples/launch_training.py:33-62``).
Observations are returned as CPU tensors because the legacy agents call
``state.numpy()`` on them (``baseline/PPO/PPO.py:209-213``).
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np
import torch

from . import _lib
from .instances import Geometry, Instance
from .vec_game import VecGame


def reward(violation, wirelength, via):
    """``-(500*violation + 4*via + 0.5*wirelength)``, evaluated exactly as
    ``/root/reference/baseline/PPO/train_PPO.py:101-102`` does."""
    r = -1
    r *= violation * 500 + via * 4 + wirelength * 0.5
    return r


def a3c_reward(openroad_cost, xroute_cost, action_list, total_step):
    """Whole-order reward of the A3C flavour, evaluated as
    ``/root/reference/baseline/A3C/utils.py:316-333`` does: cost of the default order minus cost of the
    chosen order (``0.5*wl + 4*via + 500*vio`` each, on ``[violation, wirelength, via]``), minus the
    mismatch penalty ``alpha/k * sum((a_i - i)^2)`` over the 0-based ``action_list`` with
    ``alpha = 0.1`` for the first 100 steps.  Returns ``(reward, done)``; done = no violation left."""
    cal = lambda c: 0.5 * c[1] + 4 * c[2] + 500 * c[0]
    try:
        r = cal(openroad_cost) - cal(xroute_cost)
    except Exception:
        r = 0
    alpha = 0.1 if total_step <= 100 else 0
    k = len(action_list)
    penalty = 0
    for i in range(k):
        penalty += (action_list[i] - i) ** 2
    r -= alpha / k * penalty
    done = len(xroute_cost) > 0 and xroute_cost[0] == 0
    return r, done


class Game:
    """Game wrapper (reference-compatible, GPU-backed)."""

    def __init__(self, port_recv='5556', port_initial='6667', *, geometry: Geometry | None = None,
                 instances=None, device: int = 0, max_nets: int = 64, max_aps: int = 4096, pinned_ring: int = 3):
        # the two ports are accepted for signature compatibility; there is no socket
        self.port_recv = port_recv
        self.port_initial = port_initial
        if geometry is None or instances is None:
            raise ValueError("Game needs geometry= and instances= (the region source replacing the simulator)")
        self.geometry = geometry
        self._source = itertools.cycle(instances) if isinstance(instances, (list, tuple)) else iter(instances)
        self._vec = None
        self._device = device
        self._max_nets, self._max_aps = max_nets, max_aps
        self.routed_nets = set()
        # Observations are returned as views of a small ring of pinned host buffers (device-to-host at PCIe speed, no
        # extra host copy): a returned tensor stays valid for the next pinned_ring - 1 reset/step calls, which covers
        # the agents' use (state -> numpy -> network input, baseline/PPO/PPO.py:209-213).  pinned_ring = 0 returns a
        # fresh pageable tensor per call, exactly like the reference.
        self._ring_n, self._ring, self._ring_i = pinned_ring, [], 0

    def _obs(self):
        if self._ring_n <= 0 or not torch.cuda.is_available():
            return self._vec.obs_host(0)
        need = self._vec.max_channels * self.geometry.cells
        if not self._ring or self._ring[0].numel() < need:
            self._ring = [torch.empty(need, dtype=torch.float32, pin_memory=True) for _ in range(self._ring_n)]
        self._ring_i = (self._ring_i + 1) % self._ring_n
        return self._vec.obs_host(0, out=self._ring[self._ring_i])

    def _ensure(self, inst: Instance):
        need_nets = int(inst.ap_net.max()) if len(inst.ap_net) else 1
        if self._vec is None or need_nets > self._vec.max_nets or len(inst.ap_net) > self._vec.max_aps:
            if self._vec is not None:
                self._vec.close()
            self._vec = VecGame(self.geometry, [inst], device=self._device,
                                max_nets=max(self._max_nets, need_nets),
                                max_aps=max(self._max_aps, len(inst.ap_net)))
        else:
            self._vec.load_instance(0, inst)

    def step(self, action):
        """Apply action (1-based net id).  Returns (observation, done, violation, wirelength, via)
        with the three metrics being this step's deltas (baseline_utils.py:426-433)."""
        self._vec.step(np.array([int(action)], np.int32))
        self.routed_nets.add(action)
        delta, _done, _cum = self._vec.results_host()
        violation, wirelength, via = (int(v) for v in delta[0])
        observation = self._obs()
        netSet = self._vec.legal_set(0)
        done = len(netSet) == 0
        self.legal_action_set = netSet
        return observation, done, violation, wirelength, via

    def reset(self):
        """Load the next region that has nets to route (baseline_utils.py:441-481).
        Returns (observation, reset_try_time)."""
        done = True
        reset_try_time = 0
        while done:
            inst = next(self._source)
            self._ensure(inst)
            self._vec.reset()
            self.routed_nets = set()
            self.observation = self._obs()
            self.action_space = self._vec.legal_set(0)
            if len(self.action_space) != 0:
                done = False
            else:
                reset_try_time += 1
        return self.observation, reset_try_time


def build_3Dgrid(data, routed_nets, bool_inference=False, *, device: int = 0):
    """GPU replacement of ``/root/reference/baseline/build_3Dgrid.py:224-270`` with the same
    signature and return value ``(observation, netSet, violation, wirelength, via)``.

    ``data = [[X,Y,Z], nodes, [vio, wl, via], netList]`` as produced by ``handle_messange``
    (``baseline_utils.py:9-43``).  The node list is flattened on the host; classification
    and the tensor build run behind the C ABI (``xr_build_obs_from_nodes``)."""
    L = _lib.load()
    X, Y, Z = (int(v) for v in data[0])
    nodes = data[1]
    flat = np.empty((len(nodes), 6), np.int32)
    for i, v in enumerate(nodes):
        flat[i, 0:3] = v[0]
        flat[i, 3:6] = v[2]
    max_net = int(max(flat[:, 4].max(initial=0), 0))
    keep = np.zeros(max_net + 1, np.uint8)
    if bool_inference:
        for k in data[3]:
            if 1 <= k <= max_net:
                keep[k] = 1
    else:
        keep[1:] = 1
        for k in routed_nets:
            if 1 <= int(k) <= max_net:
                keep[int(k)] = 0
    n_present = len(set(int(k) for k in flat[:, 4] if k >= 1 and keep[k]))
    out = torch.empty((1, 2 + 7 * n_present, Z, Y, X), dtype=torch.float32)
    nets = np.zeros(max(n_present, 1), np.int32)
    n_nets = C.c_int32()
    stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    rc = L.xr_build_obs_from_nodes(device, X, Y, Z, len(nodes), flat.ctypes.data_as(C.POINTER(C.c_int32)),
                                   keep.ctypes.data_as(C.POINTER(C.c_uint8)), max_net,
                                   C.cast(out.data_ptr(), C.POINTER(C.c_float)), out.numel(),
                                   nets.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n_nets), stream)
    _lib.check(rc, None)
    assert n_nets.value == n_present
    netSet = set(int(v) for v in nets[:n_nets.value])
    return out, netSet, data[2][0], data[2][1], data[2][2]
