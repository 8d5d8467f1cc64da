"""BASELINE.json configs at their full single-GPU sizes (gpu-marked): bit-exact parity on a sample of the
environments at every step, and size-independent properties on ALL of them -- the congestion metrics
recomputed from the final occupancy state (a checksum of checksums), wirelength/via monotone, every net routed."""
import numpy as np
import pytest

from xroute_env_b200.instances import make_batch, preset_geometry

pytestmark = pytest.mark.gpu


def _check_full(preset, n_envs, n_nets, sample, seed, gen_kw=None, obs_cap=-1):
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    geom = preset_geometry(preset)
    insts = make_batch(geom, n_envs, n_nets, seed, **(gen_kw or {}))
    vg = VecGame(geom, insts, device=0, obs_max_nets=obs_cap)
    vg.reset()
    rng = np.random.default_rng(seed)
    orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)     # [n_nets, n_envs]
    oracles = {e: OracleEnv(geom, insts[e]) for e in sample}
    prev = np.zeros((n_envs, 6), np.int64)
    for t in range(n_nets):
        vg.step(orders[t])
        delta, done, cum = vg.results_host()
        cum = cum.numpy().copy()
        assert (cum[:, 1] >= prev[:, 1]).all() and (cum[:, 2] >= prev[:, 2]).all()          # wirelength, via only grow
        assert (cum[:, 0] == cum[:, 3] + cum[:, 4]).all()                                    # violation = blocked + shorted
        prev = cum
        for e, orc in oracles.items():
            m = orc.step(int(orders[t, e]))
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"],
                                                 m["overflow"]], (preset, t, e)
            oc, oo, ocost = orc.last_paths(); gc, go, gcost = vg.paths(e)
            assert np.array_equal(oc, gc) and np.array_equal(ocost, gcost), (preset, t, e)
    assert bool(vg.done.all()) and int(vg.n_remaining.sum()) == 0
    # the metrics of every environment, recomputed on the host from its final occupancy
    for e in range(n_envs):
        usage, owner = vg.state(e)
        inst = insts[e]
        blk = np.zeros(usage.shape, bool)
        if len(inst.block_xyz):
            blk[inst.block_xyz[:, 2], inst.block_xyz[:, 1], inst.block_xyz[:, 0]] = True
        apn = np.zeros(usage.shape, np.int64)
        apn[inst.ap_xyz[:, 2], inst.ap_xyz[:, 1], inst.ap_xyz[:, 0]] = inst.ap_net
        blocked = int(((usage > 0) & blk).sum())
        shorted = int(((usage >= 2) | ((usage == 1) & (apn != 0) & (apn != owner))).sum())
        overflow = int(np.maximum(usage.astype(np.int64) - 1, 0).sum())
        assert [blocked, shorted, overflow] == [int(v) for v in prev[e, 3:6]], (preset, e)
    for e in sample:
        if obs_cap < 0:
            assert np.array_equal(vg.obs_host(e).numpy(), oracles[e].obs()), (preset, e)
        gu, go_ = vg.state(e); ou, oo_ = oracles[e].state()
        assert np.array_equal(gu, ou) and np.array_equal(go_, oo_), (preset, e)
    rc = vg.route_counters()
    vg.close()
    return rc


def test_config2_syn256_64_envs_full_episode():
    rc = _check_full("SYN-256", 64, 32, sample=(0, 21, 42, 63), seed=20260000)
    # hybrid policy of the default engine: the frontier search for most nets, the sweep kernels for the few-pin nets
    # with a wide bounding box (a 64-environment step does not fill the GPU with one CTA per net)
    assert rc["frontier_nets"] + rc["window_nets"] == 64 * 32 and rc["global_nets"] == 0
    assert rc["frontier_nets"] > 1500 and rc["window_nets"] > 50, rc


def test_config3_t1_7x7_shard_512_envs_full_episode():
    _check_full("T1-7x7", 512, 32, sample=(0, 100, 511), seed=31)


def test_config5_t1_1x1_shard_1024_envs_full_episode():
    _check_full("T1-1x1", 1024, 32, sample=(0, 7, 1023), seed=32, gen_kw={"max_degree": 6})


@pytest.mark.parametrize("engine", [0, 1], ids=["frontier", "sweeps"])
def test_config4_syn1024_congested_nets_on_and_off_chip(engine):
    """1024x1024x9, 128 nets clustered around hot spots (observation materialised for the first 8 nets only):
    windows up to 540x540 do not fit a cluster's shared memory, so this exercises the 16-CTA bucket and the
    full-grid HBM sweeps next to the on-chip kernels -- all bit-exact against the oracle."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    geom = preset_geometry("SYN-1024")
    insts = make_batch(geom, 2, 128, 777, hot_spots=16, hot_sigma=32.0)
    vg = VecGame(geom, insts, device=0, obs_max_nets=8, engine=engine)
    vg.reset()
    oracles = [OracleEnv(geom, i) for i in insts]
    # pick nets of very different extents: the largest and the smallest windows of each environment first
    def extent(inst, n):
        xy = inst.ap_xyz[inst.ap_net == n][:, :2]
        return int((xy.max(0) - xy.min(0)).sum())
    orders = []
    for inst in insts:
        ids = sorted(inst.net_ids, key=lambda n: -extent(inst, n))
        orders.append(ids[:3] + ids[-3:])
    for t in range(6):
        acts = np.array([o[t] for o in orders], np.int32)
        vg.step(acts)
        delta, done, cum = vg.results_host()
        for e, orc in enumerate(oracles):
            m = orc.step(int(acts[e]))
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"],
                                                 m["overflow"]], (t, e)
            oc, oo, ocost = orc.last_paths(); gc, go, gcost = vg.paths(e)
            assert np.array_equal(oc, gc) and np.array_equal(ocost, gcost), (t, e)
    rc = vg.route_counters()
    if engine == 1:
        assert rc["global_nets"] >= 1 and rc["window_nets"] >= 1, rc
    else:
        assert rc["frontier_nets"] == 12 and rc["global_nets"] == 0, rc
    vg.close()


def test_config4_syn1024_16_envs_full_episode_both_engines():
    """configs[3] at half a GPU's shard: 16 environments x 128 clustered nets on the 1024x1024x9 grid, the WHOLE episode,
    routed by the frontier engine and by the sweep engines (16-CTA windows + full-grid sweeps): cumulative metrics of every
    environment at every step and a hash of every final occupancy must agree; environment 0 is also checked bit-exactly
    (paths, costs, metrics) against the CPU oracle for the first 24 steps."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    geom = preset_geometry("SYN-1024")
    n_envs, n_nets = 16, 128
    insts = make_batch(geom, n_envs, n_nets, 4040, hot_spots=16, hot_sigma=32.0)
    rng = np.random.default_rng(4)
    orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)
    w = (np.arange(geom.cells, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(12345)) | np.uint64(1)
    out = {}
    for engine in (0, 1):
        vg = VecGame(geom, insts, device=0, obs_max_nets=8, engine=engine)
        vg.reset()
        orc = OracleEnv(geom, insts[0]) if engine == 0 else None
        cums = []
        for t in range(n_nets):
            vg.step(orders[t])
            _, _, cum = vg.results_host_np()
            cums.append(cum.copy())
            if orc is not None and t < 24:
                m = orc.step(int(orders[t, 0]))
                assert [int(v) for v in cum[0]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"], m["overflow"]], t
                oc, oo, ocost = orc.last_paths(); gc, go, gcost = vg.paths(0)
                assert np.array_equal(oc, gc) and np.array_equal(ocost, gcost), t
        assert bool(vg.done.all())
        hashes = []
        for e in range(n_envs):
            usage, owner = vg.state(e)
            hashes.append(int((usage.reshape(-1).astype(np.uint64) * w).sum() ^ (owner.reshape(-1).astype(np.uint64) * (w >> np.uint64(7))).sum()))
        out[engine] = (np.stack(cums), hashes, vg.route_counters())
        vg.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1] == out[1][1]
    rc = out[0][2]                                           # hybrid: the wide few-pin nets of this small batch take the sweep kernels
    assert rc["frontier_nets"] + rc["window_nets"] + rc["global_nets"] == n_envs * n_nets and rc["frontier_nets"] > 0
    assert out[1][2]["frontier_nets"] == 0
