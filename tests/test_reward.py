"""Reward scalarisation against the 717 known-answer tuples extracted from the reference's
shipped PPO TensorBoard log (tests/golden/reward_tfevents.npz)."""
import os

import numpy as np

from helpers import GOLD
from oracle import oracle
from xroute_env_b200.game import reward


def test_reward_matches_all_tfevents_tuples():
    t = np.load(os.path.join(GOLD, "reward_tfevents.npz"))["tuples"]
    assert t.shape == (717, 4)
    for vio, wl, via, r in t:
        assert oracle.reward(int(vio), int(wl), int(via)) == r       # exact in fp64
        assert reward(int(vio), int(wl), int(via)) == r
    assert (t[:, 0] < 0).any(), "the log holds a negative delta: deltas are signed"


def test_reference_game_episode_golden():
    """Cumulative->delta differencing, done flag and the 1-based action shift observed from
    the unmodified reference Game (tests/golden/game_episode.npz)."""
    z = np.load(os.path.join(GOLD, "game_episode.npz"))
    assert z["replies"].tolist() == [1, 0, -99]          # step(2) -> net_index 1, step(1) -> 0, b'\0' at done
    assert z["ret1"].tolist() == [0, 1, 1010, 2]
    assert z["ret2"].tolist() == [1, 0, 1000, 2]
    assert z["space0"].tolist() == [1, 2] and z["legal1"].tolist() == [1] and z["legal2"].tolist() == []
    assert z["obs0"].shape[1] == 16 and z["obs1"].shape[1] == 9 and z["obs2"].shape[1] == 2
