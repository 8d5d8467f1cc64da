"""xroute_env_b200 -- B200-native XRoute environment hot path.

Observation build, maze route of the selected net and reward/congestion metrics of the
reference's ``reset``/``step`` loop (``/root/reference/baseline/baseline_utils.py:383-481``,
``baseline/build_3Dgrid.py``) as hand-written sm_100a kernels behind a C ABI
(``include/xroute_b200.h``).  Importing the package does not load the CUDA library;
constructing ``Game``/``VecGame`` does, and raises if it is not built (no CPU fallback).
"""
from .instances import (Geometry, Instance, PRESETS, ispd18_geometry, make_batch, make_instance,
                        preset_geometry, export_data)

__all__ = ["Geometry", "Instance", "PRESETS", "ispd18_geometry", "make_batch", "make_instance",
           "preset_geometry", "export_data", "Game", "VecGame", "build_3Dgrid", "reward", "a3c_reward"]


def __getattr__(name):
    if name in ("Game", "build_3Dgrid", "reward", "a3c_reward"):
        from . import game
        return getattr(game, name)
    if name == "VecGame":
        from .vec_game import VecGame
        return VecGame
    raise AttributeError(name)
