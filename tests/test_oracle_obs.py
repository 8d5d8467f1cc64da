"""Pin the oracle's observation build to the REFERENCE: golden outputs of the reference's
own build_3Dgrid (tests/golden/obs_cases.npz, produced by tests/golden/make_golden.py),
plus a live differential run when /root/reference is mounted (build container only)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from helpers import golden_obs_cases, rows_to_data, rows_to_oracle
from xroute_env_b200.instances import export_data, ispd18_geometry, make_instance

CASES = list(golden_obs_cases())


@pytest.mark.parametrize("case", CASES, ids=[f"c{c['idx']}" for c in CASES])
def test_oracle_obs_matches_reference_golden(case):
    env, geom, inst = rows_to_oracle(case["rows"], case["dims"], routed=case["routed"],
                                     infer_netlist=case["netlist"] if case["infer"] else None)
    obs = env.obs()
    assert obs.shape == case["obs"].shape
    assert obs.dtype == np.float32
    assert np.array_equal(obs, case["obs"])
    assert env.remaining() == case["netset"]


def test_golden_quirks_present():
    """The fixtures really exercise the three bit-exactness quirks of SURVEY appendix A.3."""
    c = CASES[3 * 4 + 0] if len(CASES) > 12 else CASES[-1]
    obs = c["obs"]
    assert obs.shape[1] == 2 + 7 * len(c["netset"])
    if len(c["netset"]):
        # aliased adjacency channels: +1..+6 identical
        for r in range(len(c["netset"])):
            blk = obs[0, 2 + 7 * r: 9 + 7 * r]
            for j in range(2, 7):
                assert np.array_equal(blk[1], blk[j])
        # order channel: first n flat entries are the ascending net ids
        flat = obs[0, 1].reshape(-1)
        assert [int(v) for v in flat[: len(c["netset"])]] == c["netset"]
        assert not flat[len(c["netset"]):].any()


@pytest.mark.skipif(not os.path.exists("/root/reference/baseline/build_3Dgrid.py"),
                    reason="reference tree not mounted (GPU box)")
def test_oracle_obs_live_differential_over_an_episode():
    sys.path.insert(0, "/root/reference/baseline")
    import build_3Dgrid as ref
    from oracle.oracle import OracleEnv
    geom = ispd18_geometry(14, 11, 6)
    inst = make_instance(geom, 7, 3)
    env = OracleEnv(geom, inst)
    routed = set()
    order = np.random.default_rng(1).permutation(inst.net_ids)
    cum = (0, 0, 0)
    for net in [None] + [int(v) for v in order]:
        if net is not None:
            m = env.step(net)
            routed.add(net)
            cum = (m["violation"], m["wirelength"], m["via"])
        usage, _ = env.state()
        data = export_data(geom, inst, usage, cum)
        with contextlib.redirect_stdout(io.StringIO()):
            o, netset, v, w, a = ref.build_3Dgrid(data, routed)
        assert np.array_equal(o.numpy(), env.obs())
        assert sorted(netset) == env.remaining() and (v, w, a) == cum
