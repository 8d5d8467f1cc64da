"""gymnasium front end under the id the reference registers (``xroute_env/ordering-training-v0``,
``/root/reference/xroute_env/__init__.py:3-6``; the reference's own ``XRouteEnv`` is an empty stub,
``xroute_env/envs/core.py:1-8``, its real loop is ``baseline_utils.Game``).

``OrderingTrainingEnv``     one region at a time, gymnasium ``reset(seed, options) -> (obs, info)`` /
                            ``step(action) -> (obs, reward, terminated, truncated, info)`` on top of ``Game``
                            (actions are 1-based net ids, reward = the reference's PPO/DQN scalarisation).
``OrderingTrainingVecEnv``  the batched form on top of ``VecGame``: observations stay on the GPU (DLPack views),
                            auto-reset of finished environments like ``gymnasium.vector`` environments.

gymnasium is optional: without it the classes are plain Python objects with the same methods, and ``register()``
raises a clear error.  Nothing here touches the hot path; it is the thin API skin SURVEY.md section 7 step 2 names.
"""
from __future__ import annotations

import numpy as np

from .game import Game, reward as _reward

try:                                                   # gymnasium is not a dependency of the library
    import gymnasium as _gym
    _Base = _gym.Env
except Exception:                                      # pragma: no cover - exercised when gymnasium is absent
    _gym = None
    _Base = object

ENV_ID = "xroute_env/ordering-training-v0"


class OrderingTrainingEnv(_Base):
    """Single-region net-ordering environment (gymnasium API)."""

    metadata = {"render_modes": []}

    def __init__(self, geometry=None, instances=None, device: int = 0, **game_kw):
        self._game = Game(geometry=geometry, instances=instances, device=device, **game_kw)
        self.legal_actions = set()

    def reset(self, *, seed=None, options=None):
        obs, tries = self._game.reset()
        self.legal_actions = set(self._game.action_space)
        return obs, {"reset_try_time": tries, "legal_actions": sorted(self.legal_actions)}

    def step(self, action):
        obs, done, vio, wl, via = self._game.step(int(action))
        self.legal_actions = set(self._game.legal_action_set)
        info = {"violation": vio, "wirelength": wl, "via": via, "legal_actions": sorted(self.legal_actions)}
        return obs, float(_reward(vio, wl, via)), bool(done), False, info

    def close(self):
        if getattr(self._game, "_vec", None) is not None:
            self._game._vec.close()


class OrderingTrainingVecEnv:
    """Batched net-ordering environments on one GPU (the ``gymnasium.vector`` calling convention on ``VecGame``).

    ``reset() -> (obs, info)``; ``step(actions int32 [N]) -> (obs, reward [N] f64, terminated [N] bool, truncated [N] bool,
    info)``.  ``obs`` is the zero-copy ``[N, 2+7*max_nets, Z, Y, X]`` view of the library's buffer (valid channels of
    environment e: ``2 + 7 * n_remaining[e]``); ``info`` carries ``n_remaining``, ``legal`` (uint8 ``[N, max_nets+1]``)
    and the metric deltas, all device tensors.  Environments that finished are reset by the next ``step`` call
    (their action is ignored in that call), as gymnasium's vector environments do."""

    def __init__(self, geometry, instances, device: int = 0, **vec_kw):
        from .vec_game import VecGame
        self.vec = VecGame(geometry, instances, device=device, **vec_kw)
        self.num_envs = self.vec.n_envs
        self._needs_reset = np.zeros(self.num_envs, bool)

    def _info(self):
        return {"n_remaining": self.vec.n_remaining, "legal": self.vec.legal, "delta": self.vec.delta}

    def reset(self, *, seed=None, options=None):
        self.vec.reset()
        self._needs_reset[:] = False
        return self.vec.obs_batch(), self._info()

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int32).copy()
        if self._needs_reset.any():
            self.vec.reset(np.nonzero(self._needs_reset)[0].astype(np.int32))
            a[self._needs_reset] = 0
            self._needs_reset[:] = False
        self.vec.step(a)
        _, done, _ = self.vec.results_host_np()
        self._needs_reset = done.astype(bool) & (a != 0)
        term = self.vec.done.bool()
        return self.vec.obs_batch(), self.vec.reward, term, term.new_zeros(term.shape), self._info()

    def close(self):
        self.vec.close()


def register():
    """Register ``OrderingTrainingEnv`` with gymnasium under the reference's id."""
    if _gym is None:
        raise ImportError("gymnasium is not installed; OrderingTrainingEnv / OrderingTrainingVecEnv work without it")
    from gymnasium.envs.registration import register as _register, registry
    if ENV_ID not in registry:
        _register(id=ENV_ID, entry_point="xroute_env_b200.gym_env:OrderingTrainingEnv")
    return ENV_ID
