"""ctypes binding of libxroute_b200.so (the C ABI of include/xroute_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load the
import raises, loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``python -m xroute_env_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("XROUTE_B200_LIB", os.path.join(_HERE, "libxroute_b200.so"))   # override: A/B builds

XR_OK = 0
XR_E_INVALID, XR_E_CUDA, XR_E_ILLEGAL, XR_E_CAPACITY, XR_E_UNROUTABLE, XR_E_STATE = -1, -2, -3, -4, -5, -6
XR_M_COUNT = 6
XR_STATS_COUNT = 16
XR_K_COUNT = 9
(XR_BUF_OBS, XR_BUF_DELTA, XR_BUF_CUM, XR_BUF_DONE, XR_BUF_NREMAIN, XR_BUF_LEGAL, XR_BUF_STATS,
 XR_BUF_REWARD, XR_BUF_NETFEAT) = range(9)
K_NAMES = ["obs", "metrics", "route_begin", "sweep_xz", "sweep_y", "control", "route_win", "misc", "route_frontier"]
STAT_NAMES = ["steps", "episodes", "violation", "wirelength", "via", "blocked", "shorted", "overflow",
              "reward_x2", "relax_passes", "cells_relaxed", "connections"]


class XrConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("n_envs", C.c_int32),
        ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32),
        ("max_nets", C.c_int32), ("max_aps", C.c_int32), ("obs_max_nets", C.c_int32),
        ("path_capacity", C.c_int32),
        ("x_coords", C.POINTER(C.c_int32)), ("y_coords", C.POINTER(C.c_int32)),
        ("layer_dir", C.POINTER(C.c_uint8)),
        ("layer_pitch", C.POINTER(C.c_int32)), ("layer_min_width", C.POINTER(C.c_int32)),
        ("via_cost", C.c_int32), ("grid_cost", C.c_int32), ("drc_cost", C.c_int32),
        ("fixed_shape_cost", C.c_int32), ("block_cost", C.c_int32),
        ("pumps_per_sync", C.c_int32), ("window_margin", C.c_int32), ("min_cluster", C.c_int32),
        ("obs_mode", C.c_int32), ("engine", C.c_int32), ("metrics_mode", C.c_int32), ("guide_cost", C.c_int32), ("halo", C.c_int32),
    ]


class XrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"xroute_b200 error {code}: {msg}")
        self.code = code


class IllegalAction(XrError):
    pass


# every symbol include/xroute_b200.h declares (tests/test_abi.py checks the two agree)
SYMBOLS = [
    "xr_version", "xr_create", "xr_destroy", "xr_last_error", "xr_load_instance", "xr_load_guides", "xr_reset",
    "xr_step", "xr_step_async", "xr_step_wait", "xr_step_results", "xr_obs_layout", "xr_obs_channels", "xr_obs_copy",
    "xr_obs_dlpack", "xr_buffer_dlpack", "xr_buffer_ptr", "xr_legal_mask", "xr_get_paths",
    "xr_get_state", "xr_get_dist", "xr_stats_update", "xr_counters", "xr_profile_enable",
    "xr_profile_get", "xr_build_obs_from_nodes", "xr_route_counters", "xr_debug_counters", "xr_debug_timeline", "xr_kernel_bench",
    "xr_frontier_counters", "xr_debug_env_records", "xr_stats_allreduce",
]

_lib = None


def load():
    """Load the shared library (once).  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
    L = C.CDLL(SO_PATH)
    vp, i32p, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    i64p = C.POINTER(C.c_int64)
    L.xr_version.restype = C.c_int
    L.xr_create.restype = C.c_int
    L.xr_create.argtypes = [C.POINTER(XrConfig), C.POINTER(vp)]
    L.xr_destroy.restype = None
    L.xr_destroy.argtypes = [vp]
    L.xr_last_error.restype = C.c_char_p
    L.xr_last_error.argtypes = [vp]
    L.xr_load_instance.restype = C.c_int
    L.xr_load_instance.argtypes = [vp, C.c_int32, C.c_int32, i32p, C.c_int32, i32p, i32p, i32p]
    L.xr_load_guides.restype = C.c_int
    L.xr_load_guides.argtypes = [vp, C.c_int32, C.c_int32, i32p]
    L.xr_reset.restype = C.c_int
    L.xr_reset.argtypes = [vp, i32p, C.c_int32, vp]
    L.xr_step.restype = C.c_int
    L.xr_step.argtypes = [vp, i32p, vp]
    L.xr_step_async.restype = C.c_int
    L.xr_step_async.argtypes = [vp, i32p, vp]
    L.xr_step_wait.restype = C.c_int
    L.xr_step_wait.argtypes = [vp]
    L.xr_step_results.restype = C.c_int
    L.xr_step_results.argtypes = [vp, i32p, u8p, i64p, vp]
    L.xr_obs_layout.restype = C.c_int
    L.xr_obs_layout.argtypes = [vp, i64p, i32p]
    L.xr_obs_channels.restype = C.c_int
    L.xr_obs_channels.argtypes = [vp, C.c_int32, i32p]
    L.xr_obs_copy.restype = C.c_int
    L.xr_obs_copy.argtypes = [vp, C.c_int32, C.POINTER(C.c_float), C.c_int64, vp]
    L.xr_obs_dlpack.restype = C.c_int
    L.xr_obs_dlpack.argtypes = [vp, C.c_int32, C.POINTER(vp)]
    L.xr_buffer_dlpack.restype = C.c_int
    L.xr_buffer_dlpack.argtypes = [vp, C.c_int32, C.POINTER(vp)]
    L.xr_buffer_ptr.restype = C.c_int
    L.xr_buffer_ptr.argtypes = [vp, C.c_int32, C.POINTER(vp), i64p]
    L.xr_legal_mask.restype = C.c_int
    L.xr_legal_mask.argtypes = [vp, C.c_int32, u8p, i32p]
    L.xr_get_paths.restype = C.c_int
    L.xr_get_paths.argtypes = [vp, C.c_int32, i32p, C.c_int32, i32p, i32p, C.POINTER(C.c_uint32), C.c_int32, i32p]
    L.xr_get_state.restype = C.c_int
    L.xr_get_state.argtypes = [vp, C.c_int32, u8p, C.POINTER(C.c_uint16)]
    L.xr_get_dist.restype = C.c_int
    L.xr_get_dist.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint32)]
    L.xr_stats_update.restype = C.c_int
    L.xr_stats_update.argtypes = [vp, vp]
    L.xr_stats_allreduce.restype = C.c_int
    L.xr_stats_allreduce.argtypes = [vp, vp, vp]
    L.xr_counters.restype = C.c_int
    L.xr_counters.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.xr_route_counters.restype = C.c_int
    L.xr_route_counters.argtypes = [vp, i64p, i64p, i64p]
    L.xr_debug_env_records.restype = C.c_int
    L.xr_debug_env_records.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.xr_frontier_counters.restype = C.c_int
    L.xr_frontier_counters.argtypes = [vp, i64p, i64p]
    L.xr_kernel_bench.restype = C.c_int
    L.xr_kernel_bench.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]
    L.xr_debug_timeline.restype = C.c_int
    L.xr_debug_timeline.argtypes = [vp, C.POINTER(C.c_double)]
    L.xr_debug_counters.restype = C.c_int
    L.xr_debug_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.xr_profile_enable.restype = C.c_int
    L.xr_profile_enable.argtypes = [vp, C.c_int32]
    L.xr_profile_get.restype = C.c_int
    L.xr_profile_get.argtypes = [vp, C.POINTER(C.c_double), i64p]
    L.xr_build_obs_from_nodes.restype = C.c_int
    L.xr_build_obs_from_nodes.argtypes = [C.c_int32] * 5 + [i32p, u8p, C.c_int32, C.POINTER(C.c_float),
                                                              C.c_int64, i32p, i32p, vp]
    _lib = L
    return L


def check(rc: int, handle=None):
    if rc == XR_OK:
        return
    msg = load().xr_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc == XR_E_ILLEGAL:
        raise IllegalAction(rc, msg)
    raise XrError(rc, msg)


# PyCapsule plumbing for DLPack: torch.from_dlpack consumes a capsule named "dltensor".
_PyCapsule_New = C.pythonapi.PyCapsule_New
_PyCapsule_New.restype = C.py_object
_PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]


def capsule(managed_tensor_ptr: int):
    return _PyCapsule_New(managed_tensor_ptr, b"dltensor", None)
