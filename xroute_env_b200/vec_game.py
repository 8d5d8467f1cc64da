"""VecGame: a batch of routing environments on one B200, stepped in lockstep.

The batched counterpart of the reference's ``Game``
(``/root/reference/baseline/baseline_utils.py:383-481``): the same per-environment
``reset``/``step`` semantics (1-based net-id actions, cumulative-metric differencing,
``done`` when no net remains, ``[1, 2+7n, Z, Y, X]`` float32 observations in the
``build_3Dgrid`` layout) for N independent regions at once, with every output left
on the GPU and handed to torch zero-copy through DLPack.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .instances import Geometry, Instance


def _i32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class VecGame:
    def __init__(self, geom: Geometry, instances: list[Instance], *, device: int = 0,
                 max_nets: int | None = None, max_aps: int | None = None,
                 obs_max_nets: int = -1, path_capacity: int = 0, pumps_per_sync: int = 0,
                 window_margin: int = 0, min_cluster: int = 0, obs_mode: int = 0, engine: int = 0,
                 metrics_mode: int = 0, guide_cost: int = 0, halo: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.geom = geom
        self.insts = list(instances)
        self.n_envs = len(instances)
        self.device = device
        if max_nets is None:
            max_nets = max([int(i.ap_net.max()) if len(i.ap_net) else 1 for i in instances] + [1])
        if max_aps is None:
            max_aps = max([len(i.ap_net) for i in instances] + [1])
        self.max_nets, self.max_aps = max_nets, max_aps
        self._keep = [np.ascontiguousarray(geom.x_coords, np.int32), np.ascontiguousarray(geom.y_coords, np.int32),
                      np.ascontiguousarray(geom.layer_dir, np.uint8),
                      np.ascontiguousarray(geom.layer_pitch, np.int32),
                      np.ascontiguousarray(geom.layer_min_width, np.int32)]
        cfg = _lib.XrConfig()
        cfg.device, cfg.n_envs = device, self.n_envs
        cfg.X, cfg.Y, cfg.Z = geom.X, geom.Y, geom.Z
        cfg.max_nets, cfg.max_aps, cfg.obs_max_nets = max_nets, max_aps, obs_max_nets
        cfg.path_capacity = path_capacity
        cfg.x_coords, cfg.y_coords = _i32p(self._keep[0]), _i32p(self._keep[1])
        cfg.layer_dir = self._keep[2].ctypes.data_as(C.POINTER(C.c_uint8))
        cfg.layer_pitch, cfg.layer_min_width = _i32p(self._keep[3]), _i32p(self._keep[4])
        cfg.via_cost, cfg.grid_cost, cfg.drc_cost = geom.via_cost, geom.grid_cost, geom.drc_cost
        cfg.fixed_shape_cost, cfg.block_cost = geom.fixed_shape_cost, geom.block_cost
        cfg.pumps_per_sync = pumps_per_sync
        cfg.window_margin, cfg.min_cluster = window_margin, min_cluster
        cfg.obs_mode = obs_mode
        cfg.engine, cfg.metrics_mode = engine, metrics_mode
        cfg.guide_cost, cfg.halo = guide_cost, halo
        self.guide_cost = guide_cost
        rc = self._L.xr_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            _lib.check(rc, None)
        for i, inst in enumerate(instances):
            self.load_instance(i, inst)
        stride, maxc = C.c_int64(), C.c_int32()
        self._L.xr_obs_layout(self._h, C.byref(stride), C.byref(maxc))
        self.obs_stride, self.max_channels = stride.value, maxc.value
        # pinned host staging for the e2e path
        pin = torch.cuda.is_available()
        self._h_delta = torch.zeros((self.n_envs, 3), dtype=torch.int32, pin_memory=pin)
        self._h_done = torch.zeros((self.n_envs,), dtype=torch.uint8, pin_memory=pin)
        self._h_cum = torch.zeros((self.n_envs, _lib.XR_M_COUNT), dtype=torch.int64, pin_memory=pin)
        self._views = {}
        self._pin_obs = None
        self._h_np = (self._h_delta.numpy(), self._h_done.numpy(), self._h_cum.numpy())   # same pinned memory

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._views = {}
            self._L.xr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_instance(self, env_id: int, inst: Instance):
        self.insts[env_id] = inst
        b = np.ascontiguousarray(inst.block_xyz, np.int32).reshape(-1)
        n = np.ascontiguousarray(inst.ap_net, np.int32)
        p = np.ascontiguousarray(inst.ap_pin, np.int32)
        x = np.ascontiguousarray(inst.ap_xyz, np.int32).reshape(-1)
        _lib.check(self._L.xr_load_instance(self._h, env_id, len(b) // 3, _i32p(b), len(n), _i32p(n), _i32p(p),
                                            _i32p(x)), self._h)
        if self.guide_cost > 0:                      # route guides of the instance (optional cost term)
            gd = getattr(inst, "guides", None)
            gb = np.ascontiguousarray(np.zeros((0, 6), np.int32) if gd is None else gd, np.int32).reshape(-1, 6)
            _lib.check(self._L.xr_load_guides(self._h, env_id, len(gb), _i32p(gb)), self._h)

    # ----------------------------------------------------------------- reset/step
    def reset(self, env_ids=None):
        """Reset all (or the listed) environments; observations are rebuilt on device."""
        if env_ids is None:
            rc = self._L.xr_reset(self._h, None, 0, self._stream())
        else:
            ids = np.ascontiguousarray(env_ids, np.int32)
            rc = self._L.xr_reset(self._h, _i32p(ids), len(ids), self._stream())
        _lib.check(rc, self._h)

    def step(self, actions):
        """actions: host int32 [N] (numpy / list / CPU tensor): >=1 net id, 0 idle, -1 stop.
        Routes the chosen nets; results stay on the GPU (see the ``delta``/``done``/... views)."""
        if isinstance(actions, torch.Tensor):
            actions = actions.cpu().numpy()
        a = np.ascontiguousarray(actions, np.int32)
        if a.shape != (self.n_envs,):
            raise ValueError(f"actions must have shape ({self.n_envs},)")
        _lib.check(self._L.xr_step(self._h, _i32p(a), self._stream()), self._h)

    def step_async(self, actions):
        """Enqueue a step without waiting for the device (``xr_step_async``); ``step_wait`` (or any call that reads
        state) completes it.  Host work placed between the two overlaps the routing kernels."""
        if isinstance(actions, torch.Tensor):
            actions = actions.cpu().numpy()
        a = np.ascontiguousarray(actions, np.int32)
        if a.shape != (self.n_envs,):
            raise ValueError(f"actions must have shape ({self.n_envs},)")
        _lib.check(self._L.xr_step_async(self._h, _i32p(a), self._stream()), self._h)

    def step_wait(self):
        _lib.check(self._L.xr_step_wait(self._h), self._h)

    def results_host(self):
        """(delta int32 [N,3], done uint8 [N], cum int64 [N,6]) copied to pinned host memory."""
        _lib.check(self._L.xr_step_results(
            self._h, C.cast(self._h_delta.data_ptr(), C.POINTER(C.c_int32)),
            C.cast(self._h_done.data_ptr(), C.POINTER(C.c_uint8)),
            C.cast(self._h_cum.data_ptr(), C.POINTER(C.c_int64)), self._stream()), self._h)
        return self._h_delta, self._h_done, self._h_cum

    def results_host_np(self):
        """``results_host`` as numpy views of the same pinned buffers (no tensor-op overhead on the host)."""
        self.results_host()
        return self._h_np

    # -------------------------------------------------------------- zero-copy views
    def _buffer(self, which: int) -> torch.Tensor:
        if which not in self._views:
            out = C.c_void_p()
            _lib.check(self._L.xr_buffer_dlpack(self._h, which, C.byref(out)), self._h)
            self._views[which] = torch.from_dlpack(_lib.capsule(out.value))
        return self._views[which]

    @property
    def delta(self) -> torch.Tensor:       # int32 [N,3] d_violation, d_wirelength, d_via
        return self._buffer(_lib.XR_BUF_DELTA)

    @property
    def cum(self) -> torch.Tensor:         # int64 [N,6]
        return self._buffer(_lib.XR_BUF_CUM)

    @property
    def done(self) -> torch.Tensor:        # uint8 [N]
        return self._buffer(_lib.XR_BUF_DONE)

    @property
    def n_remaining(self) -> torch.Tensor:  # int32 [N]
        return self._buffer(_lib.XR_BUF_NREMAIN)

    @property
    def legal(self) -> torch.Tensor:       # uint8 [N, max_nets+1]
        return self._buffer(_lib.XR_BUF_LEGAL)

    @property
    def reward(self) -> torch.Tensor:      # float64 [N]
        return self._buffer(_lib.XR_BUF_REWARD)

    @property
    def net_features(self) -> torch.Tensor:
        """float32 [N, max_nets+1, 22] per-net feature vectors (row = net id) in the layout of the
        reference's A3C flavour (``baseline/A3C/utils.py:212-277``): half-perimeter of the AP box,
        nets with an AP inside it, 16 layer flags, times routed since reset, last d_violation /
        d_wirelength / d_via.  Static part set at load, dynamic part updated by every step."""
        return self._buffer(_lib.XR_BUF_NETFEAT)

    def route_order(self, orders) -> torch.Tensor:
        """Whole-order action of the A3C / MCTS flavours (``Response.net_list``,
        ``baseline/xroute/net_ordering.proto:78``): reset, then route ``orders[k]`` (int32 [n, N],
        0 = idle) as n back-to-back batched steps without reading anything back.  Returns the
        cumulative metrics int64 [N, 6] view; the cost of an order is ``order_cost(cum)``."""
        o = np.ascontiguousarray(orders, np.int32)
        if o.ndim != 2 or o.shape[1] != self.n_envs:
            raise ValueError(f"orders must have shape (n, {self.n_envs})")
        self.reset()
        for k in range(o.shape[0]):
            self.step(o[k])
        return self.cum

    @staticmethod
    def order_cost(cum: torch.Tensor) -> torch.Tensor:
        """``0.5*wirelength + 4*via + 500*violation`` (``baseline/A3C/utils.py:195-196``), float64 [N]."""
        c = cum.to(torch.float64)
        return 0.5 * c[:, 1] + 4.0 * c[:, 2] + 500.0 * c[:, 0]

    def obs_batch(self) -> torch.Tensor:
        """float32 [N, max_channels, Z, Y, X] strided view of the whole observation block
        (channels beyond 2+7*n_remaining[e] of environment e are stale)."""
        if "obs" not in self._views:
            out = C.c_void_p()
            _lib.check(self._L.xr_obs_dlpack(self._h, -1, C.byref(out)), self._h)
            self._views["obs"] = torch.from_dlpack(_lib.capsule(out.value))
        return self._views["obs"]

    def obs(self, env_id: int) -> torch.Tensor:
        """float32 CUDA [1, 2+7n, Z, Y, X] view of one environment (zero-copy)."""
        out = C.c_void_p()
        _lib.check(self._L.xr_obs_dlpack(self._h, env_id, C.byref(out)), self._h)
        return torch.from_dlpack(_lib.capsule(out.value))

    def obs_host(self, env_id: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """CPU copy of one environment's observation (what the reference's Game returns), float32 [1, 2+7n, Z, Y, X].

        ``out``: a caller-owned CPU float32 tensor with at least ``(2+7n) * cells`` elements, ideally pinned
        (``torch.empty(..., pin_memory=True)``): the device-to-host copy goes straight into it at PCIe speed and the
        result is a view of it.  Without ``out`` the copy is staged through an internal pinned buffer and cloned into a
        fresh pageable tensor the caller owns (one extra host memcpy)."""
        ch = C.c_int32()
        _lib.check(self._L.xr_obs_channels(self._h, env_id, C.byref(ch)), self._h)
        g = self.geom
        n = ch.value * g.cells
        if out is not None:
            if out.dtype != torch.float32 or out.device.type != "cpu" or not out.is_contiguous() or out.numel() < n:
                raise ValueError("out must be a contiguous CPU float32 tensor with room for the observation")
            _lib.check(self._L.xr_obs_copy(self._h, env_id, C.cast(out.data_ptr(), C.POINTER(C.c_float)), out.numel(),
                                           self._stream()), self._h)
            return out.view(-1)[:n].view(1, ch.value, g.Z, g.Y, g.X)
        # through a pinned staging buffer (a pageable destination makes the driver stage the copy itself,
        # several times slower), then one host memcpy into the fresh tensor the caller owns
        if self._pin_obs is None and torch.cuda.is_available() and self.max_channels * g.cells * 4 <= (2 << 30):
            self._pin_obs = torch.empty(self.max_channels * g.cells, dtype=torch.float32, pin_memory=True)
        if self._pin_obs is None:
            out = torch.empty((1, ch.value, g.Z, g.Y, g.X), dtype=torch.float32)
            _lib.check(self._L.xr_obs_copy(self._h, env_id, C.cast(out.data_ptr(), C.POINTER(C.c_float)),
                                           out.numel(), self._stream()), self._h)
            return out
        _lib.check(self._L.xr_obs_copy(self._h, env_id, C.cast(self._pin_obs.data_ptr(), C.POINTER(C.c_float)),
                                       n, self._stream()), self._h)
        return self._pin_obs[:n].clone().view(1, ch.value, g.Z, g.Y, g.X)

    # ------------------------------------------------------------------- host info
    def legal_set(self, env_id: int) -> set[int]:
        mask = np.zeros(self.max_nets + 1, np.uint8)
        n = C.c_int32()
        _lib.check(self._L.xr_legal_mask(self._h, env_id, mask.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n)),
                   self._h)
        return set(int(i) for i in np.nonzero(mask)[0])

    def paths(self, env_id: int):
        """Last routed net's paths: (cells int32, conn_off int32 [k+1], conn_cost uint32 [k])."""
        n_cells, n_conn = C.c_int32(), C.c_int32()
        _lib.check(self._L.xr_get_paths(self._h, env_id, None, 0, C.byref(n_cells), None, None, 0, C.byref(n_conn)),
                   self._h)
        cells = np.zeros(max(n_cells.value, 1), np.int32)
        off = np.zeros(n_conn.value + 1, np.int32)
        cost = np.zeros(max(n_conn.value, 1), np.uint32)
        _lib.check(self._L.xr_get_paths(self._h, env_id, _i32p(cells), len(cells), C.byref(n_cells), _i32p(off),
                                        cost.ctypes.data_as(C.POINTER(C.c_uint32)), len(cost), C.byref(n_conn)),
                   self._h)
        return cells[:n_cells.value], off, cost[:n_conn.value]

    def state(self, env_id: int):
        g = self.geom
        usage = np.zeros(g.cells, np.uint8)
        owner = np.zeros(g.cells, np.uint16)
        _lib.check(self._L.xr_get_state(self._h, env_id, usage.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        owner.ctypes.data_as(C.POINTER(C.c_uint16))), self._h)
        return usage.reshape(g.Z, g.Y, g.X), owner.reshape(g.Z, g.Y, g.X)

    def dist(self, env_id: int) -> np.ndarray:
        g = self.geom
        d = np.zeros(g.cells, np.uint32)
        _lib.check(self._L.xr_get_dist(self._h, env_id, d.ctypes.data_as(C.POINTER(C.c_uint32))), self._h)
        return d.reshape(g.Z, g.Y, g.X)

    def stats(self) -> torch.Tensor:
        """int64 [16] per-handle sums on the GPU (all-reduce with SUM across ranks)."""
        _lib.check(self._L.xr_stats_update(self._h, self._stream()), self._h)
        return self._buffer(_lib.XR_BUF_STATS)

    def stats_allreduce(self, nccl_comm: int) -> torch.Tensor:
        """``stats()`` summed in place over the ranks of an ``ncclComm_t`` (given as an integer address) by the library
        itself (``xr_stats_allreduce``): the multi-GPU path of a consumer that does not use ``torch.distributed``."""
        _lib.check(self._L.xr_stats_allreduce(self._h, C.c_void_p(nccl_comm), self._stream()), self._h)
        return self._buffer(_lib.XR_BUF_STATS)

    def counters(self) -> dict:
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._L.xr_counters(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), self._h)
        return {"kernel_launches": a.value, "relax_passes": b.value, "cells_relaxed": c.value,
                "host_syncs": d.value}

    def route_counters(self) -> dict:
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._L.xr_route_counters(self._h, C.byref(a), C.byref(b), C.byref(c)), self._h)
        f, r = C.c_int64(), C.c_int64()
        _lib.check(self._L.xr_frontier_counters(self._h, C.byref(f), C.byref(r)), self._h)
        return {"window_nets": a.value, "global_nets": b.value, "window_fallbacks": c.value, "frontier_nets": f.value}

    def debug_counters(self) -> dict:
        out = (C.c_uint64 * 16)()
        _lib.check(self._L.xr_debug_counters(self._h, out), self._h)
        names = ["win_iterations", "win_connections", "win_relax_cycles", "win_kernel_cycles", "win_nets", "win_area",
                 "c8_iterations", "c8_relax_cycles", "ph_compact_y", "ph_sweep_y", "ph_compact_x", "ph_sweep_x",
                 "ph_sweep_z", "ph_halo_pull", "ph_end_sync", "c8_connections"]
        return {k: int(out[i]) for i, k in enumerate(names) if k}

    def kernel_bench(self, which: str, reps: int = 10) -> dict:
        """Isolated timing of 'obs' or 'metrics' over all environments (roofline accounting)."""
        ms, nbytes = C.c_double(), C.c_double()
        _lib.check(self._L.xr_kernel_bench(self._h, _lib.K_NAMES.index(which), reps, C.byref(ms), C.byref(nbytes),
                                           self._stream()), self._h)
        return {"ms": ms.value, "bytes": nbytes.value, "gbs": nbytes.value / (ms.value * 1e-3) / 1e9 if ms.value else 0.0}

    def debug_timeline(self) -> dict:
        out = (C.c_double * 9)()
        _lib.check(self._L.xr_debug_timeline(self._h, out), self._h)
        return {f"g{g}_{n}": round(out[3 * g + k], 3) for g in range(3)
                for k, n in enumerate(("route_start", "route_end", "obs_end"))}

    def profile(self, enable: bool):
        _lib.check(self._L.xr_profile_enable(self._h, int(enable)), self._h)

    def profile_get(self) -> dict:
        ms = (C.c_double * _lib.XR_K_COUNT)()
        n = (C.c_int64 * _lib.XR_K_COUNT)()
        _lib.check(self._L.xr_profile_get(self._h, ms, n), self._h)
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(_lib.K_NAMES)}
