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


def test_a3c_whole_order_reward_formula():
    """baseline/A3C/utils.py:316-333 on hand-computed cases."""
    from xroute_env_b200.game import a3c_reward
    # default order cost 0.5*2000+4*3+500*1 = 1512, chosen order 0.5*1800+4*2 = 908; order [1,0,2]: penalty (1-0)^2+(0-1)^2 = 2
    r, done = a3c_reward([1, 2000, 3], [0, 1800, 2], [1, 0, 2], total_step=5)
    assert r == 1512 - 908 - 0.1 / 3 * 2 and done
    r, done = a3c_reward([1, 2000, 3], [2, 1800, 2], [0, 1, 2], total_step=500)
    assert r == 1512 - (908 + 1000) and not done
    r, done = a3c_reward([1, 2000, 3], [], [0], total_step=500)       # malformed cost: the reference falls back to 0
    assert r == 0 and not done
