"""ctypes binding of the CPU oracle (oracle/xr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(xroute_env_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libxr_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "xr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        p = C.c_void_p
        i32p = C.POINTER(C.c_int32)
        L.orc_create.restype = p
        L.orc_create.argtypes = [C.c_int] * 3 + [i32p, i32p, C.POINTER(C.c_uint8), i32p, i32p] + [C.c_int] * 5
        L.orc_destroy.argtypes = [p]
        L.orc_load.restype = C.c_int
        L.orc_load.argtypes = [p, C.c_int, i32p, C.c_int, i32p, i32p, i32p]
        L.orc_reset.argtypes = [p]
        L.orc_set_options.argtypes = [p, C.c_int, C.c_int]
        L.orc_load_guides.argtypes = [p, C.c_int, i32p]
        L.orc_remaining.restype = C.c_int
        L.orc_remaining.argtypes = [p, i32p]
        L.orc_obs.restype = C.c_int
        L.orc_obs.argtypes = [p, C.POINTER(C.c_float), C.c_int]
        L.orc_export_nodes.argtypes = [p, i32p]
        L.orc_step.restype = C.c_int
        L.orc_step.argtypes = [p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_path_len.restype = C.c_int64
        L.orc_path_len.argtypes = [p]
        L.orc_conn_count.restype = C.c_int
        L.orc_conn_count.argtypes = [p]
        L.orc_get_path.argtypes = [p, i32p, i32p, C.POINTER(C.c_uint32)]
        L.orc_settled.restype = C.c_int64
        L.orc_settled.argtypes = [p]
        L.orc_src_pin.restype = C.c_int
        L.orc_src_pin.argtypes = [p, C.c_int]
        L.orc_get_state.argtypes = [p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)]
        L.orc_distance_field.restype = C.c_int
        L.orc_distance_field.argtypes = [p, C.c_int, i32p, C.c_int, C.POINTER(C.c_uint32)]
        L.orc_set_usage.argtypes = [p, C.POINTER(C.c_uint8)]
        L.orc_set_routed.restype = C.c_int
        L.orc_set_routed.argtypes = [p, C.c_int, C.c_int]
        L.orc_reward.restype = C.c_double
        L.orc_reward.argtypes = [C.c_int64] * 3
        _lib = L
    return _lib


def _i32(a):
    a = np.ascontiguousarray(a, np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def reward(violation: int, wirelength: int, via: int) -> float:
    return float(lib().orc_reward(int(violation), int(wirelength), int(via)))


class OracleEnv:
    """Single-region CPU environment with the oracle's route/commit/metrics/obs."""

    def __init__(self, geom, inst=None, *, guide_cost: int = 0, halo: int = 0):
        L = lib()
        self.geom = geom
        xc, xcp = _i32(geom.x_coords)
        yc, ycp = _i32(geom.y_coords)
        ld = np.ascontiguousarray(geom.layer_dir, np.uint8)
        pi, pip_ = _i32(geom.layer_pitch)
        mw, mwp = _i32(geom.layer_min_width)
        self._h = L.orc_create(geom.X, geom.Y, geom.Z, xcp, ycp, ld.ctypes.data_as(C.POINTER(C.c_uint8)),
                               pip_, mwp, geom.via_cost, geom.grid_cost, geom.drc_cost,
                               geom.fixed_shape_cost, geom.block_cost)
        self.inst = None
        L.orc_set_options(self._h, int(guide_cost), int(halo))
        if inst is not None:
            self.load(inst)

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def load(self, inst):
        b, bp = _i32(inst.block_xyz.reshape(-1))
        n, np_ = _i32(inst.ap_net)
        p, pp = _i32(inst.ap_pin)
        x, xp = _i32(inst.ap_xyz.reshape(-1))
        rc = lib().orc_load(self._h, len(inst.block_xyz), bp, len(inst.ap_net), np_, pp, xp)
        if rc != 0:
            raise ValueError(f"orc_load failed: {rc}")
        g = getattr(inst, "guides", None)
        gb, gbp = _i32(np.zeros((0, 6), np.int32) if g is None else np.asarray(g, np.int32).reshape(-1, 6))
        lib().orc_load_guides(self._h, len(gb), gbp)
        self.inst = inst
        self.reset()

    def reset(self):
        lib().orc_reset(self._h)

    def remaining(self) -> list[int]:
        n = lib().orc_remaining(self._h, None)
        out = np.zeros(max(n, 1), np.int32)
        lib().orc_remaining(self._h, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return [int(v) for v in out[:n]]

    def obs(self) -> np.ndarray:
        """float32 [1, 2+7n, Z, Y, X] exactly as build_3Dgrid returns it."""
        g = self.geom
        n = len(self.remaining())
        Cn = 2 + 7 * n
        out = np.empty((Cn, g.cells), np.float32)
        rc = lib().orc_obs(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), Cn)
        assert rc == Cn, rc
        return out.reshape(1, Cn, g.Z, g.Y, g.X)

    def step(self, net: int, full: bool = False) -> dict:
        out = np.zeros(10, np.int64)
        rc = lib().orc_step(self._h, int(net), int(full), out.ctypes.data_as(C.POINTER(C.c_int64)))
        if rc != 0:
            raise ValueError(f"orc_step({net}) failed: {rc}")
        keys = ["d_violation", "d_wirelength", "d_via", "violation", "wirelength", "via",
                "blocked", "shorted", "overflow", "done"]
        return {k: int(v) for k, v in zip(keys, out)}

    def last_paths(self):
        """(cells int32 [n], conn_off int32 [k+1], conn_cost uint32 [k]) of the last step;
        cells are canonical indices (z*Y + y)*X + x, target first."""
        L = lib()
        n = L.orc_path_len(self._h)
        k = L.orc_conn_count(self._h)
        cells = np.zeros(max(n, 1), np.int32)
        off = np.zeros(k + 1, np.int32)
        cost = np.zeros(max(k, 1), np.uint32)
        L.orc_get_path(self._h, cells.ctypes.data_as(C.POINTER(C.c_int32)),
                       off.ctypes.data_as(C.POINTER(C.c_int32)),
                       cost.ctypes.data_as(C.POINTER(C.c_uint32)))
        return cells[:n], off, cost[:k]

    def settled(self) -> int:
        return int(lib().orc_settled(self._h))

    def src_pin(self, net: int) -> int:
        return int(lib().orc_src_pin(self._h, net))

    def state(self):
        g = self.geom
        usage = np.zeros(g.cells, np.uint8)
        owner = np.zeros(g.cells, np.uint16)
        lib().orc_get_state(self._h, usage.ctypes.data_as(C.POINTER(C.c_uint8)),
                            owner.ctypes.data_as(C.POINTER(C.c_uint16)))
        return usage.reshape(g.Z, g.Y, g.X), owner.reshape(g.Z, g.Y, g.X)

    def set_usage(self, usage: np.ndarray):
        u = np.ascontiguousarray(usage, np.uint8).reshape(-1)
        assert u.size == self.geom.cells
        lib().orc_set_usage(self._h, u.ctypes.data_as(C.POINTER(C.c_uint8)))

    def set_routed(self, net: int, flag: bool = True):
        lib().orc_set_routed(self._h, int(net), int(flag))

    def distance_field(self, net: int, src_cells) -> np.ndarray:
        g = self.geom
        s, sp = _i32(src_cells)
        out = np.zeros(g.cells, np.uint32)
        rc = lib().orc_distance_field(self._h, int(net), sp, len(s), out.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert rc == 0
        return out.reshape(g.Z, g.Y, g.X)


def net_features(geom, inst, count=None, delta=None) -> np.ndarray:
    """CPU restatement of the A3C flavour's 22-feature per-net vectors
    (/root/reference/baseline/A3C/utils.py:212-277), float64 [max_net+1, 22], row = net id:
    [0] half-perimeter of the access points' box in point coordinates (x, y DBU, z = layer; :246),
    [1] number of nets -- the net itself included, the loop at :250-256 never skips it -- with an
    access point inside that box, [2..17] flag per layer (maze z) holding an access point (:259-262),
    [18] count_map entry (:268), [19..21] metrics_delta entry (:270).  Test infrastructure only."""
    nets = inst.net_ids
    n_max = max(nets) if nets else 0
    out = np.zeros((n_max + 1, 22), np.float64)
    px = geom.x_coords.astype(np.int64)[inst.ap_xyz[:, 0]] if len(inst.ap_xyz) else np.zeros(0, np.int64)
    py = geom.y_coords.astype(np.int64)[inst.ap_xyz[:, 1]] if len(inst.ap_xyz) else np.zeros(0, np.int64)
    pz = inst.ap_xyz[:, 2].astype(np.int64) if len(inst.ap_xyz) else np.zeros(0, np.int64)
    for n in nets:
        m = inst.ap_net == n
        x0, x1, y0, y1, z0, z1 = px[m].min(), px[m].max(), py[m].min(), py[m].max(), pz[m].min(), pz[m].max()
        out[n, 0] = (x1 - x0) + (y1 - y0) + (z1 - z0)
        inside = (px >= x0) & (px <= x1) & (py >= y0) & (py <= y1) & (pz >= z0) & (pz <= z1)
        out[n, 1] = len(set(inst.ap_net[inside].tolist()))
        for z in set(pz[m].tolist()):
            if z < 16:
                out[n, 2 + z] = 1
        if count is not None:
            out[n, 18] = count.get(n, 0)
        if delta is not None:
            out[n, 19:22] = delta.get(n, (0, 0, 0))
    return out


def order_cost(violation, wirelength, via) -> float:
    """/root/reference/baseline/A3C/utils.py:195-196."""
    return 0.5 * wirelength + 4 * via + 500 * violation
