"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded instances -- paths, connection costs, occupancy state, metrics and observations
must be BIT-EXACT (integer / index work; the float32 observation holds only 0/1 and
small integers)."""
import numpy as np
import pytest

from xroute_env_b200.instances import Geometry, ispd18_geometry, make_batch, make_instance

pytestmark = pytest.mark.gpu


def _run_episode(geom, insts, seed=0, check_obs_every=1, steps=None, **vgkw):
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    vg = VecGame(geom, insts, device=0, **vgkw)
    vg.reset()
    oracles = [OracleEnv(geom, i, guide_cost=vgkw.get("guide_cost", 0), halo=vgkw.get("halo", 0)) for i in insts]
    rng = np.random.default_rng(seed)
    orders = [list(rng.permutation(i.net_ids)) for i in insts]
    for e, orc in enumerate(oracles):
        assert np.array_equal(vg.obs_host(e).numpy(), orc.obs()), f"initial obs env {e}"
        assert vg.legal_set(e) == set(orc.remaining())
    T = max(len(o) for o in orders)
    if steps is not None:
        T = min(T, steps)
    for t in range(T):
        acts = np.array([int(o[t]) if t < len(o) else 0 for o in orders], np.int32)
        vg.step(acts)
        delta, done, cum = vg.results_host()
        rew = vg.reward.cpu().numpy()
        for e, orc in enumerate(oracles):
            if acts[e] == 0:
                assert [int(v) for v in delta[e]] == [0, 0, 0]
                continue
            m = orc.step(int(acts[e]))
            oc, oo, ocost = orc.last_paths()
            gc, go, gcost = vg.paths(e)
            assert np.array_equal(ocost, gcost), (t, e, ocost, gcost)
            assert np.array_equal(oo, go), (t, e)
            assert np.array_equal(oc, gc), (t, e)
            assert [int(v) for v in delta[e]] == [m["d_violation"], m["d_wirelength"], m["d_via"]], (t, e)
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"],
                                                 m["shorted"], m["overflow"]], (t, e)
            assert int(done[e]) == m["done"]
            assert rew[e] == -(500 * m["d_violation"] + 4 * m["d_via"] + 0.5 * m["d_wirelength"])
            if t % check_obs_every == 0 or m["done"]:
                assert np.array_equal(vg.obs_host(e).numpy(), orc.obs()), (t, e)
                gu, go_ = vg.state(e)
                ou, oo_ = orc.state()
                assert np.array_equal(gu, ou) and np.array_equal(go_, oo_), (t, e)
            assert vg.legal_set(e) == set(orc.remaining())
    vg.close()


@pytest.mark.parametrize("shape", [(25, 26, 9), (12, 10, 5), (40, 36, 9), (33, 31, 3), (64, 64, 2), (70, 9, 9)])
def test_episode_small_grids(shape):
    geom = ispd18_geometry(*shape)
    insts = make_batch(geom, 6, 8, seed=100 + shape[0])
    _run_episode(geom, insts, seed=shape[1])


@pytest.mark.parametrize("kw", [dict(window_margin=-1), dict(window_margin=1), dict(window_margin=3, min_cluster=2),
                                dict(min_cluster=4), dict(min_cluster=8), dict(min_cluster=16), dict(window_margin=30)],
                         ids=["global-only", "margin1-fallbacks", "margin3-c2", "c4", "c8", "c16", "margin30"])
def test_route_paths_agree_across_engines(kw):
    """Window kernel (every cluster size), forced fall-backs to the full-grid sweeps and the
    full-grid path alone must all reproduce the oracle bit-exactly."""
    geom = ispd18_geometry(48, 44, 9)
    insts = make_batch(geom, 5, 10, seed=900, p_obstacle=0.2)
    _run_episode(geom, insts, seed=8, engine=1, **kw)


@pytest.mark.parametrize("knobs", [{}, {"XR_FR_RAY": "1"}, {"XR_FR_RAY": "3", "XR_FR_DELTA": "0"}, {"XR_FR_DELTA": "400"},
                                   {"XR_FR_DELTA": "1000000"}, {"XR_FR_CAP": "64"}, {"XR_FR_THREADS": "64", "XR_FR_CAP": "128"},
                                   {"XR_FR_THREADS": "128", "XR_FR_RAY": "5"}, {"XR_FR_PARK": "1", "XR_FR_BAND": "1"},
                                   {"XR_FR_PARK": "16", "XR_FR_DELTA": "100", "XR_FR_DMAX": "1", "XR_FR_CAP": "64"}, {"XR_FR_PARK": "0"}],
                         ids=["default", "ray1", "ray3-delta0", "delta400", "delta-inf", "spill", "t64-spill", "t128-ray5",
                              "park1", "park16-spill", "nopark"])
def test_frontier_engine_knobs(knobs, monkeypatch):
    """The default engine (goal-directed frontier search, csrc/xr_frontier.cu) for every ray length, bucket width,
    block size, with open lists that spill to global memory and with far-list parking from the first entry on: the knobs change the schedule of the relaxations,
    never a distance the target choice or the walk reads -- paths, costs, metrics and observations stay bit-exact.
    metrics_mode 1 on top checks the commit-maintained congestion counts against the full scan."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    geom = ispd18_geometry(48, 44, 9)
    insts = make_batch(geom, 5, 10, seed=900, p_obstacle=0.2)
    _run_episode(geom, insts, seed=8)
    _run_episode(geom, insts[:2], seed=9, metrics_mode=1)
    geom = ispd18_geometry(90, 70, 9)
    geom.x_coords = np.cumsum(np.random.default_rng(1).integers(100, 700, 90)).astype(np.int32)   # non-uniform pitch
    _run_episode(geom, make_batch(geom, 3, 8, seed=901), seed=10, check_obs_every=4)


@pytest.mark.parametrize("metrics_mode", [0, 1])
def test_sweeps_beside_frontier_one_step(metrics_mode, monkeypatch):
    """One step that uses every route path at once: nets on the frontier kernel, nets on window clusters (some of which
    escape their window and are handed over), and nets whose window fits no cluster (XR_WIN_FIT_CAP shrinks the fit
    limit) -- those are pumped by the full-grid sweeps on their own stream while the other kernels run
    (xr_step_wait).  Bit-exact against the oracle, with both metric modes."""
    monkeypatch.setenv("XR_HYBRID_AREA", "30")
    monkeypatch.setenv("XR_HYBRID_PINS", "30")
    monkeypatch.setenv("XR_WIN_FIT_CAP", "9000")
    geom = ispd18_geometry(90, 70, 9)
    insts = make_batch(geom, 12, 10, seed=4100, p_obstacle=0.15, max_degree=5)
    from xroute_env_b200 import VecGame
    vg = VecGame(geom, insts, device=0, window_margin=1)
    vg.reset()
    rng = np.random.default_rng(3)
    orders = [list(rng.permutation(i.net_ids)) for i in insts]
    for t in range(10):
        vg.step(np.array([int(o[t]) for o in orders], np.int32))
    c = vg.route_counters()
    vg.close()
    assert c["frontier_nets"] > 0 and c["window_nets"] > 0 and c["global_nets"] > 0 and c["window_fallbacks"] > 0, c
    _run_episode(geom, insts, seed=3, window_margin=1, metrics_mode=metrics_mode, check_obs_every=3)


def _with_guides(geom, insts, margin, layers):
    """Synthetic route guides: per net the bounding box of its access points grown by `margin` cells, on `layers`."""
    out = []
    for inst in insts:
        boxes = []
        for n in inst.net_ids:
            xy = inst.ap_xyz[inst.ap_net == n]
            x0, y0 = xy[:, 0].min() - margin, xy[:, 1].min() - margin
            x1, y1 = xy[:, 0].max() + margin, xy[:, 1].max() + margin
            if n % 5 == 0:
                continue                                   # some nets come without guides: no guide term for them
            for z in layers:
                boxes.append((n, x0, x1, y0, y1, z))
        inst.guides = np.array(boxes, np.int32).reshape(-1, 6)
        out.append(inst)
    return out


@pytest.mark.parametrize("kw", [dict(guide_cost=1), dict(guide_cost=4, halo=1), dict(halo=2), dict(guide_cost=1, halo=1, metrics_mode=1)],
                         ids=["guide1", "guide4-halo1", "halo2", "guide1-halo1-scan"])
def test_optional_cost_terms_guide_and_halo(kw):
    """The two optional terms of the pinned run configuration (-follow_guide 1 / GUIDECOST, SHAPEBLOATWIDTH): out-of-guide
    multiplier and the spacing halo of routed wires, oracle and frontier engine together, bit-exact on congested
    instances where both change the routes (checked: the same episodes route differently with the terms off)."""
    from oracle.oracle import OracleEnv
    geom = ispd18_geometry(40, 36, 6)
    insts = _with_guides(geom, make_batch(geom, 4, 12, seed=321, p_obstacle=0.2), margin=1, layers=(0, 2))
    _run_episode(geom, insts, seed=4, check_obs_every=4, **kw)
    # the terms are not inert: with them off at least one route of the episode differs
    a, b = OracleEnv(geom, insts[0]), OracleEnv(geom, insts[0], guide_cost=kw.get("guide_cost", 0), halo=kw.get("halo", 0))
    differs = False
    for net in insts[0].net_ids:
        a.step(net); b.step(net)
        differs |= not np.array_equal(a.last_paths()[0], b.last_paths()[0])
    assert differs
    geom = ispd18_geometry(70, 50, 9)
    geom.x_coords = np.cumsum(np.random.default_rng(2).integers(100, 700, 70)).astype(np.int32)
    insts = _with_guides(geom, make_batch(geom, 3, 10, seed=322), margin=0, layers=(0, 1, 2, 3))
    _run_episode(geom, insts, seed=5, check_obs_every=5, **kw)


def test_optional_cost_terms_need_the_frontier_engine():
    from xroute_env_b200 import VecGame
    geom = ispd18_geometry(20, 20, 3)
    with pytest.raises(RuntimeError, match="frontier engine"):
        VecGame(geom, make_batch(geom, 1, 3, seed=1), device=0, engine=1, halo=1)


@pytest.mark.parametrize("engine", [0, 1], ids=["frontier", "sweeps"])
def test_metrics_by_scan_and_by_commit_agree(engine):
    """Congested instance (shorts, blocked cells, overflow all non-zero): the counts the commits maintain
    (metrics_mode 0) and the per-step scan of the occupancy field (metrics_mode 1) both equal the oracle's."""
    geom = ispd18_geometry(30, 30, 3)
    insts = make_batch(geom, 4, 24, seed=41, p_obstacle=0.35)
    for mm in (0, 1):
        _run_episode(geom, insts, seed=6, engine=engine, metrics_mode=mm, check_obs_every=6)


@pytest.mark.parametrize("obs_mode", [0, 1], ids=["incremental", "full-rebuild"])
def test_observation_modes_match_oracle(obs_mode):
    """In-place incremental observation update (default) and the full per-step rebuild both
    reproduce the reference layout bit-exactly at every step, across resets and capped nets."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    geom = ispd18_geometry(30, 28, 9)
    insts = make_batch(geom, 4, 9, seed=930)
    _run_episode(geom, insts, seed=10, obs_mode=obs_mode)
    # second episode on the same handle, interrupted by a partial reset
    vg = VecGame(geom, insts, device=0, obs_mode=obs_mode)
    oracles = [OracleEnv(geom, i) for i in insts]
    for rep in range(2):
        vg.reset()
        for o in oracles:
            o.reset()
        for t, net in enumerate((5, 2, 9, 1)):
            vg.step(np.array([net] * 4, np.int32))
            for e, o in enumerate(oracles):
                o.step(net)
                assert np.array_equal(vg.obs_host(e).numpy(), o.obs()), (rep, t, e)
        vg.reset([0, 2])
        oracles[0].reset(); oracles[2].reset()
        vg.step(np.array([3, 3, 3, 3], np.int32))
        for e, o in enumerate(oracles):
            o.step(3)
            assert np.array_equal(vg.obs_host(e).numpy(), o.obs()), (rep, "after partial reset", e)
    # a new instance in slot 1 invalidates its buffer: the next reset rebuilds it
    other = make_instance(geom, 5, 931)
    vg.load_instance(1, other)
    vg.reset([1])
    assert np.array_equal(vg.obs_host(1).numpy(), OracleEnv(geom, other).obs())
    vg.close()


@pytest.mark.parametrize("minc,cluster", [(2, 0), (2, 4), (8, 0), (16, 0)], ids=["dual-c2", "dual-c4", "dual-c8", "dual-c16"])
def test_dual_cyclic_layout_kernel(minc, cluster, monkeypatch):
    """Every net through the dual-layout window kernel (rows and columns dealt cyclically over the
    cluster, lowered cells pushed between the two copies over DSMEM): bit-exact like the others,
    including forced window escapes that hand a connection to the full-grid path."""
    monkeypatch.setenv("XR_DUAL_PINS", "2")
    monkeypatch.setenv("XR_DUAL_MINC", str(minc))
    geom = ispd18_geometry(48, 44, 9)
    insts = make_batch(geom, 5, 10, seed=900, p_obstacle=0.2)
    _run_episode(geom, insts, seed=8, min_cluster=cluster, engine=1)
    _run_episode(geom, insts[:2], seed=9, min_cluster=cluster, window_margin=1, engine=1)
    geom = ispd18_geometry(90, 70, 9)
    geom.x_coords = np.cumsum(np.random.default_rng(1).integers(100, 700, 90)).astype(np.int32)   # non-uniform pitch
    _run_episode(geom, make_batch(geom, 3, 8, seed=901), seed=10, min_cluster=cluster, check_obs_every=4, engine=1)


@pytest.mark.parametrize("seed", range(4))
def test_tie_heavy_grids(seed, monkeypatch):
    """Grids made of ties (`tools/fuzz_parity.py ... ties`, fixed seeds): every pitch and via cost a small multiple of
    100, random layer directions and cost constants -- most cells have several equal-cost predecessors, so the paths
    rest entirely on the canonical target / backtrace rules.  Odd seeds force every net through the dual kernel."""
    rng = np.random.default_rng(500 + seed)
    X, Y, Z = int(rng.integers(20, 70)), int(rng.integers(20, 70)), int(rng.integers(2, 10))
    geom = ispd18_geometry(X, Y, Z)
    px, py = rng.choice([100, 200, 300], 2)
    geom.x_coords = (np.cumsum(rng.choice([1, 1, 1, 2, 3], X)) * px).astype(np.int32)
    geom.y_coords = (py * np.arange(Y)).astype(np.int32)
    geom.layer_dir = rng.integers(0, 2, Z).astype(np.uint8)
    geom.layer_pitch = rng.choice([25, 50, 75, 100], Z).astype(np.int32)
    geom.layer_min_width = rng.choice([5, 10, 20], Z).astype(np.int32)
    geom.via_cost, geom.grid_cost = int(rng.choice([1, 2, 4])), int(rng.choice([0, 1, 2]))
    geom.drc_cost, geom.fixed_shape_cost, geom.block_cost = int(rng.choice([1, 2, 8])), int(rng.choice([1, 2, 8])), int(rng.choice([1, 5, 32]))
    if seed % 2:
        monkeypatch.setenv("XR_DUAL_PINS", "2"); monkeypatch.setenv("XR_DUAL_MINC", "2")
    insts = make_batch(geom, 4, 8, seed=510 + seed, p_obstacle=0.25)
    _run_episode(geom, insts, seed=seed, check_obs_every=3)                                        # frontier engine
    _run_episode(geom, insts, seed=seed, min_cluster=2 if seed < 2 else 0, check_obs_every=3, engine=1)
    _run_episode(geom, insts[:2], seed=seed, window_margin=-1, check_obs_every=8, engine=1)


def _strip_instances():
    """Three 700 x 10 x 3 strips with the generator's blockages and six hand-placed nets whose pins sit at the two ends
    (and, for the 3- and 4-pin nets, in the middle) of the strip."""
    from xroute_env_b200.instances import Instance
    geom = ispd18_geometry(700, 10, 3)
    insts = []
    for e, base in enumerate(make_batch(geom, 3, 6, seed=77, p_obstacle=0.15)):
        aps = []
        for k in range(1, 7):
            aps += [(k, 1, 5 + 3 * k + e, (k + e) % 10, 0), (k, 2, 694 - 2 * k - e, (3 * k + e) % 10, k % 2)]
            if k >= 4:
                aps += [(k, 3, 300 + 17 * k, (k + 2 * e) % 10, 1), (k, 3, 301 + 17 * k, (k + 2 * e) % 10, 1)]
            if k == 6:
                aps += [(k, 4, 500 + e, 2, 0)]
        a = np.array(aps, np.int32)
        taken = set(map(tuple, a[:, 2:5].tolist()))
        blk = np.array([b for b in base.block_xyz.tolist() if tuple(b) not in taken], np.int32).reshape(-1, 3)
        insts.append(Instance(block_xyz=blk, ap_net=a[:, 0].copy(), ap_pin=a[:, 1].copy(), ap_xyz=np.ascontiguousarray(a[:, 2:5])))
    return geom, insts


@pytest.mark.parametrize("engine", ["band", "dual", "frontier"])
def test_long_paths_span_several_commit_flushes(engine, monkeypatch):
    """A 700-track-wide strip: single connections of several hundred cells.  The window kernels stage at most 256 path
    cells on chip before a parallel commit pass, and put the whole walk on the tree from the path record afterwards --
    paths longer than one stage must come out bit-exact too, and stay on the window engines."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    if engine == "dual":
        monkeypatch.setenv("XR_DUAL_PINS", "2"); monkeypatch.setenv("XR_DUAL_MINC", "2")
    geom, insts = _strip_instances()
    vg = VecGame(geom, insts, device=0, min_cluster=2, engine=0 if engine == "frontier" else 1)
    vg.reset()
    orcs = [OracleEnv(geom, i) for i in insts]
    longest = 0
    for net in range(1, 7):
        vg.step(np.array([net] * 3, np.int32))
        _, _, cum = vg.results_host()
        for e, o in enumerate(orcs):
            m = o.step(net)
            oc, oo, ocost = o.last_paths(); gc, go, gcost = vg.paths(e)
            assert np.array_equal(ocost, gcost) and np.array_equal(oo, go) and np.array_equal(oc, gc), (net, e)
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"], m["overflow"]]
            longest = max([longest] + np.diff(oo).tolist())
    for e, o in enumerate(orcs):
        assert np.array_equal(vg.obs_host(e).numpy(), o.obs())
        assert np.array_equal(vg.state(e)[0], o.state()[0])
    rc = vg.route_counters()
    vg.close()
    assert longest > 300, longest
    if engine != "frontier":
        assert rc["window_nets"] > 0 and rc["global_nets"] == 0, rc


def _own_walk_case():
    """Layer 0 is horizontal with x pitch 300 and y pitch 100, so an x step (300) costs exactly a y step
    (100 x (1 + GRIDCOST)).  Net 1: the walk back from (2,6,0) runs along row 6 to (5,6,0), whose only real predecessor is the
    source (5,5,0) below it -- but the cell it just left, (4,6,0), is one x step away too.  Net 2: the same on a later
    connection (three pins: the walk from (7,11,0) meets the tree at (9,10,0) from (9,11,0)); net 3: a control."""
    from xroute_env_b200.instances import Instance
    geom = ispd18_geometry(14, 16, 3)
    geom.x_coords = (300 * np.arange(14)).astype(np.int32)
    geom.y_coords = (100 * np.arange(16)).astype(np.int32)
    aps = [(1, 1, 5, 5, 0), (1, 2, 2, 6, 0),
           (2, 1, 9, 10, 0), (2, 2, 12, 10, 0), (2, 3, 7, 11, 0),
           (3, 1, 3, 12, 0), (3, 2, 2, 9, 0)]
    a = np.array(aps, np.int32)
    inst = Instance(block_xyz=np.zeros((0, 3), np.int32), ap_net=a[:, 0].copy(), ap_pin=a[:, 1].copy(),
                    ap_xyz=np.ascontiguousarray(a[:, 2:5]))
    return geom, inst


@pytest.mark.parametrize("engine", ["band", "dual", "global", "frontier"])
def test_backtrace_ignores_the_cells_of_its_own_walk(engine, monkeypatch):
    """Found by tools/fuzz_parity.py (non-uniform grid, configuration 762 of seed 7): every engine used to turn the
    cells of a walk into sources (distance 0) while still walking, and a later cell of the same walk could then accept
    the cell it had just left as predecessor (0 + w == dist) ahead of the real source in the canonical order."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    geom, inst = _own_walk_case()
    kw = dict(min_cluster=2)
    if engine == "dual":
        monkeypatch.setenv("XR_DUAL_PINS", "2"); monkeypatch.setenv("XR_DUAL_MINC", "2")
    if engine == "global":
        kw = dict(window_margin=-1)
    vg = VecGame(geom, [inst], device=0, engine=0 if engine == "frontier" else 1, **kw)
    vg.reset()
    orc = OracleEnv(geom, inst)
    for net in (1, 2, 3):
        vg.step(np.array([net], np.int32))
        orc.step(net)
        oc, oo, ocost = orc.last_paths()
        gc, go, gcost = vg.paths(0)
        assert np.array_equal(ocost, gcost) and np.array_equal(oo, go), net
        assert np.array_equal(oc, gc), (net, oc.tolist(), gc.tolist())
        assert len(set(gc[go[0]:go[1]].tolist())) == go[1] - go[0], "a path visits a cell twice"
    assert np.array_equal(vg.state(0)[0], orc.state()[0])
    vg.close()


def test_path_capacity_overflow_fails_the_step():
    """The router reads a net's new tree cells back from the path record: a record too small for the net is an error
    of the step (XR_E_CAPACITY), never a silently different route."""
    from xroute_env_b200 import VecGame
    geom = ispd18_geometry(40, 40, 3)
    insts = make_batch(geom, 1, 4, seed=5)
    vg = VecGame(geom, insts, device=0, path_capacity=3)
    vg.reset()
    with pytest.raises(RuntimeError, match="path_capacity"):
        for net in insts[0].net_ids:
            vg.step(np.array([net], np.int32))
    # the failed step left the batch half-stepped: the handle refuses to step until the environments are reset
    from xroute_env_b200._lib import XrError, XR_E_STATE
    with pytest.raises(XrError) as ei:
        vg.step(np.array([insts[0].net_ids[-1]], np.int32))
    assert ei.value.code == XR_E_STATE
    vg.reset()
    assert vg.legal_set(0) == set(insts[0].net_ids)
    with pytest.raises(RuntimeError, match="path_capacity"):
        for net in insts[0].net_ids:
            vg.step(np.array([net], np.int32))
    vg.close()


def test_step_async_and_wait():
    """xr_step_async enqueues a step without blocking; xr_step_wait (or any call that reads state) completes it; a second
    step before that is a call-sequence error.  Results equal the synchronous path's."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    from xroute_env_b200._lib import XrError, XR_E_STATE
    geom = ispd18_geometry(40, 36, 9)
    insts = make_batch(geom, 4, 6, seed=7)
    vg = VecGame(geom, insts, device=0)
    vg.reset()
    orcs = [OracleEnv(geom, i) for i in insts]
    order = insts[0].net_ids
    for k, net in enumerate(order):
        acts = np.array([net] * 4, np.int32)
        vg.step_async(acts)
        acts[:] = 0                                             # the array may be reused at once
        if k == 1:
            with pytest.raises(XrError) as ei:
                vg.step_async(np.array([order[-1]] * 4, np.int32))
            assert ei.value.code == XR_E_STATE
        if k % 2 == 0:
            vg.step_wait()
        delta, done, cum = vg.results_host()                    # completes the pending step when nobody waited
        for e, o in enumerate(orcs):
            m = o.step(net)
            assert [int(v) for v in cum[e]] == [m["violation"], m["wirelength"], m["via"], m["blocked"], m["shorted"], m["overflow"]]
            assert np.array_equal(vg.obs_host(e).numpy(), o.obs())
    vg.close()


def test_window_fallback_counter_and_exactness():
    from xroute_env_b200 import VecGame
    geom = ispd18_geometry(64, 64, 9)
    insts = make_batch(geom, 4, 12, seed=910, p_obstacle=0.3)
    _run_episode(geom, insts, seed=9, window_margin=1, engine=1)
    vg = VecGame(geom, insts, device=0, window_margin=1, engine=1)
    vg.reset()
    for net in (1, 2, 3, 4, 5, 6):
        vg.step(np.array([net] * 4, np.int32))
    rc = vg.route_counters()
    assert rc["window_nets"] == 24 and rc["global_nets"] == 0
    assert rc["window_fallbacks"] >= 1, "margin 1 with 30% blockage must trip the exit test somewhere"
    vg.close()


def test_episode_t1_7x7():
    geom = ispd18_geometry(112, 116, 9)
    insts = make_batch(geom, 4, 16, seed=5)
    _run_episode(geom, insts, seed=1, check_obs_every=5)


def test_episode_syn256_partial():
    geom = ispd18_geometry(256, 256, 9)
    insts = make_batch(geom, 2, 12, seed=11)
    _run_episode(geom, insts, seed=2, check_obs_every=6, steps=12)


def test_wide_grid_x300():
    # X not a multiple of the warp tile: exercises padded rows / masked lanes
    geom = ispd18_geometry(300, 40, 4)
    insts = make_batch(geom, 2, 6, seed=21)
    _run_episode(geom, insts, seed=3)


def test_tall_grid_y600():
    geom = ispd18_geometry(40, 600, 3)
    insts = make_batch(geom, 2, 6, seed=22)
    _run_episode(geom, insts, seed=4)


def test_nonuniform_tracks():
    rng = np.random.default_rng(9)
    X, Y, Z = 45, 38, 6
    geom = ispd18_geometry(X, Y, Z)
    geom.x_coords = np.cumsum(rng.integers(100, 700, X)).astype(np.int32)
    geom.y_coords = np.cumsum(rng.integers(100, 700, Y)).astype(np.int32)
    insts = make_batch(geom, 4, 8, seed=31)
    _run_episode(geom, insts, seed=5)


def test_congested_many_pins():
    # dense pins, heavy blockage: shorts / blocked cells / overflow become non-zero
    geom = ispd18_geometry(30, 30, 3)
    insts = make_batch(geom, 4, 24, seed=41, p_obstacle=0.35)
    _run_episode(geom, insts, seed=6)


def test_ragged_batch_and_idle_actions():
    # environments with different numbers of nets: finished ones get action 0
    geom = ispd18_geometry(32, 32, 4)
    insts = [make_instance(geom, n, 50 + n) for n in (1, 3, 7, 5)]
    _run_episode(geom, insts, seed=7)


def test_illegal_action_leaves_state_unchanged():
    from xroute_env_b200 import VecGame
    from xroute_env_b200._lib import IllegalAction
    geom = ispd18_geometry(25, 26, 9)
    insts = make_batch(geom, 2, 4, seed=61)
    vg = VecGame(geom, insts, device=0)
    vg.reset()
    before = vg.obs_host(0).clone()
    with pytest.raises(IllegalAction):
        vg.step(np.array([99, 1], np.int32))
    vg.step(np.array([1, 1], np.int32))
    with pytest.raises(IllegalAction):
        vg.step(np.array([1, 2], np.int32))          # net 1 already routed in env 0
    assert vg.legal_set(0) == {2, 3, 4}
    vg.reset()
    assert np.array_equal(vg.obs_host(0).numpy(), before.numpy())
    vg.close()


def test_reset_subset_and_second_episode():
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    geom = ispd18_geometry(28, 28, 9)
    insts = make_batch(geom, 3, 5, seed=71)
    vg = VecGame(geom, insts, device=0)
    vg.reset()
    oracles = [OracleEnv(geom, i) for i in insts]
    for net in (2, 4):
        vg.step(np.array([net] * 3, np.int32))
        for o in oracles:
            o.step(net)
    vg.reset([1])
    oracles[1].reset()
    for e in range(3):
        assert np.array_equal(vg.obs_host(e).numpy(), oracles[e].obs())
    vg.step(np.array([1, 1, 1], np.int32))
    delta, done, cum = vg.results_host()
    for e in range(3):
        m = oracles[e].step(1)
        assert [int(v) for v in cum[e]][:3] == [m["violation"], m["wirelength"], m["via"]]
        assert np.array_equal(vg.obs_host(e).numpy(), oracles[e].obs())
    vg.close()


def test_dlpack_views_alias_device_memory():
    import torch
    from xroute_env_b200 import VecGame
    geom = ispd18_geometry(25, 26, 9)
    insts = make_batch(geom, 3, 4, seed=81)
    vg = VecGame(geom, insts, device=0)
    vg.reset()
    ob = vg.obs_batch()
    assert ob.is_cuda and ob.dtype == torch.float32 and ob.shape == (3, 2 + 7 * 4, 9, 26, 25)
    o1 = vg.obs(1)
    assert o1.shape == (1, 30, 9, 26, 25) and o1.data_ptr() == ob[1].data_ptr()
    assert torch.equal(o1.cpu(), vg.obs_host(1))
    vg.step(np.array([1, 2, 3], np.int32))
    assert vg.obs(1).shape == (1, 23, 9, 26, 25)
    assert vg.n_remaining.cpu().tolist() == [3, 3, 3]
    assert vg.legal.cpu()[1].tolist() == [0, 1, 0, 1, 1]
    s = vg.stats().cpu()
    assert s[0] == 3 and s[1] == 0
    vg.close()


def test_build_3dgrid_dropin_matches_oracle():
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import build_3Dgrid
    from xroute_env_b200.instances import export_data
    geom = ispd18_geometry(14, 12, 5)
    inst = make_instance(geom, 6, 91)
    orc = OracleEnv(geom, inst)
    routed = set()
    for net in (3, 5):
        m = orc.step(net)
        routed.add(net)
    usage, _ = orc.state()
    data = export_data(geom, inst, usage, (m["violation"], m["wirelength"], m["via"]))
    obs, nets, v, w, a = build_3Dgrid(data, routed)
    assert np.array_equal(obs.numpy(), orc.obs())
    assert nets == set(orc.remaining()) and (v, w, a) == (m["violation"], m["wirelength"], m["via"])
    obs2, nets2, *_ = build_3Dgrid(data, set(), bool_inference=True)
    assert nets2 == set(inst.net_ids) and obs2.shape[1] == 2 + 7 * 6


def test_build_3dgrid_dropin_matches_reference_golden():
    """GPU build_3Dgrid drop-in against outputs of the reference's own build_3Dgrid."""
    from helpers import golden_obs_cases, rows_to_data
    from xroute_env_b200 import build_3Dgrid
    n = 0
    for c in golden_obs_cases():
        data = rows_to_data(c["rows"], c["dims"], (3, 1010, 2), c["netlist"])
        obs, nets, v, w, a = build_3Dgrid(data, set(c["routed"]), bool_inference=c["infer"])
        assert np.array_equal(obs.numpy(), c["obs"]), c["idx"]
        assert sorted(nets) == c["netset"] and (v, w, a) == (3, 1010, 2)
        n += 1
    assert n >= 20


def test_game_dropin_episode_matches_oracle():
    """The reference-compatible Game (reset/step signatures of baseline_utils.py:383-481)."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import Game
    from xroute_env_b200.instances import Instance
    geom = ispd18_geometry(25, 26, 9)
    empty = Instance(block_xyz=np.zeros((0, 3), np.int32), ap_net=np.zeros(0, np.int32),
                     ap_pin=np.zeros(0, np.int32), ap_xyz=np.zeros((0, 3), np.int32))
    inst = make_instance(geom, 5, 77)
    game = Game(geometry=geom, instances=[empty, inst])
    obs, tries = game.reset()
    assert tries == 1                                    # the empty region is skipped and counted
    orc = OracleEnv(geom, inst)
    assert obs.device.type == "cpu" and np.array_equal(obs.numpy(), orc.obs())
    assert game.action_space == set(inst.net_ids)
    for net in (3, 1, 5, 2, 4):
        obs, done, vio, wl, via = game.step(net)
        m = orc.step(net)
        assert (vio, wl, via) == (m["d_violation"], m["d_wirelength"], m["d_via"]) and done == bool(m["done"])
        assert np.array_equal(obs.numpy(), orc.obs()) and game.legal_action_set == set(orc.remaining())
    assert done and obs.shape[1] == 2 and game.routed_nets == {1, 2, 3, 4, 5}


def test_ppo_rollout_example_runs():
    """configs[4]: policy consumes the DLPack observation block; whole episodes complete."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "ppo_rollout.py"), "--envs", "64", "--nets", "6",
                          "--steps", "14"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "env-steps/s" in out.stdout and "episodes finished 128" in out.stdout, out.stdout


@pytest.mark.parametrize("name", ["t1_7x7_y79800", "t1_7x7_y319200", "t1_7x7_y79800_union", "t1_1x1_gx3_gy6",
                                  "t1_1x1_gx2_gy6"])
def test_real_ispd18_test1_regions(name):
    """Row f1 / configs[0]: the regions extracted from the ispd18_test1 LEF/DEF/guide files
    (tests/golden/ispd18_test1_regions.npz, tools/make_ispd_regions.py) -- two copies of the
    region in one batch routed in different random orders, bit-exact against the oracle.
    The *_union region lies on the union of all layers' tracks (non-uniform pitch)."""
    import os
    from xroute_env_b200.ispd import load_regions
    g, inst = load_regions(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                        "ispd18_test1_regions.npz"))[name]
    _run_episode(g, [inst, inst], seed=17, check_obs_every=5)


@pytest.mark.parametrize("name", ["t1_7x7_y79800", "t1_7x7_y319200", "t1_1x1_gx3_gy6"])
def test_real_regions_with_the_pinned_guide_and_halo_terms(name):
    """The real ispd18_test1 regions with their real route guides (ispd18_test1.input.guide clipped to the region) under
    the run configuration the reference pins (-follow_guide 1, GUIDECOST 1, SHAPEBLOATWIDTH -> one track of halo):
    frontier engine == oracle, bit-exact."""
    import os
    from xroute_env_b200.ispd import load_regions
    g, inst = load_regions(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                        "ispd18_test1_regions.npz"))[name]
    assert inst.guides is not None and len(inst.guides) > 0
    _run_episode(g, [inst, inst], seed=23, check_obs_every=6, guide_cost=1, halo=1)


def test_stop_idle_single_pin_and_state_errors():
    """Protocol edge cases of net_ordering.proto: -1 stops an environment (:48), 0 leaves it untouched, a net with
    a single pin is only marked routed (nothing to connect), stepping before the first reset is an error."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    from xroute_env_b200._lib import XrError
    from xroute_env_b200.instances import Instance
    geom = ispd18_geometry(16, 14, 4)
    inst = Instance(block_xyz=np.array([[5, 5, 0], [6, 5, 0]], np.int32),
                    ap_net=np.array([1, 1, 2, 2, 3, 3, 3], np.int32), ap_pin=np.array([1, 1, 1, 2, 1, 2, 3], np.int32),
                    ap_xyz=np.array([[2, 2, 0], [3, 2, 0], [1, 8, 0], [12, 9, 1], [4, 11, 0], [9, 3, 1], [13, 12, 0]], np.int32))
    vg = VecGame(geom, [inst, inst, inst], device=0)
    with pytest.raises(XrError):
        vg.step(np.array([1, 1, 1], np.int32))                      # before reset
    vg.reset()
    orc = [OracleEnv(geom, inst) for _ in range(3)]
    obs1_before = vg.obs_host(1).clone()
    vg.step(np.array([1, 0, -1], np.int32))                         # single-pin net | idle | stop
    delta, done, cum = vg.results_host()
    m = orc[0].step(1)
    assert [int(v) for v in delta[0]] == [m["d_violation"], m["d_wirelength"], m["d_via"]] == [0, 0, 0]
    assert vg.legal_set(0) == {2, 3} == set(orc[0].remaining()) and not int(done[0])
    assert np.array_equal(vg.obs_host(0).numpy(), orc[0].obs())
    assert [int(v) for v in delta[1]] == [0, 0, 0] and vg.legal_set(1) == {1, 2, 3} and not int(done[1])
    assert np.array_equal(vg.obs_host(1).numpy(), obs1_before.numpy())
    assert int(done[2]) == 1 and [int(v) for v in delta[2]] == [0, 0, 0]
    assert vg.legal.cpu()[2].sum().item() == 0 and vg.legal_set(2) == set()      # a stopped environment offers no action
    with pytest.raises(XrError):
        vg.step(np.array([0, 0, 2], np.int32))                      # environment 2 is stopped
    vg.step(np.array([3, 2, 0], np.int32))
    delta, done, cum = vg.results_host()
    for e, net in ((0, 3), (1, 2)):
        m = orc[e].step(net)
        assert [int(v) for v in cum[e]][:3] == [m["violation"], m["wirelength"], m["via"]]
        oc, oo, ocost = orc[e].last_paths(); gc, go, gcost = vg.paths(e)
        assert np.array_equal(oc, gc) and np.array_equal(ocost, gcost)
        assert np.array_equal(vg.obs_host(e).numpy(), orc[e].obs())
    vg.reset([2])
    assert vg.legal_set(2) == {1, 2, 3} and not int(vg.done.cpu()[2])
    vg.close()


def test_gym_style_front_ends():
    """gymnasium calling convention on Game / VecGame (xroute_env_b200.gym_env): 5-tuple steps, rewards of the reference's
    scalarisation, auto-reset of finished environments in the vector form."""
    import torch
    from oracle.oracle import OracleEnv
    from xroute_env_b200.gym_env import OrderingTrainingEnv, OrderingTrainingVecEnv
    geom = ispd18_geometry(25, 26, 9)
    insts = make_batch(geom, 3, 4, seed=61)
    env = OrderingTrainingEnv(geometry=geom, instances=[insts[0]])
    obs, info = env.reset(seed=0)
    orc = OracleEnv(geom, insts[0])
    assert info["legal_actions"] == insts[0].net_ids and np.array_equal(obs.numpy(), orc.obs())
    for k, net in enumerate(insts[0].net_ids):
        obs, rew, term, trunc, info = env.step(net)
        m = orc.step(net)
        assert rew == -(500 * m["d_violation"] + 4 * m["d_via"] + 0.5 * m["d_wirelength"]) and not trunc
        assert term == (k == 3) and np.array_equal(obs.numpy(), orc.obs())
    env.close()
    venv = OrderingTrainingVecEnv(geom, insts)
    obs, info = venv.reset()
    assert obs.is_cuda and obs.shape[0] == 3 and info["n_remaining"].tolist() == [4, 4, 4]
    for net in (1, 2, 3, 4):
        obs, rew, term, trunc, info = venv.step(np.array([net] * 3, np.int32))
    assert term.all() and not trunc.any()
    obs, rew, term, trunc, info = venv.step(np.array([1, 1, 1], np.int32))      # finished environments are reset, action ignored
    assert info["n_remaining"].tolist() == [4, 4, 4] and not term.any() and float(rew.abs().sum()) == 0.0
    venv.close()


def test_batched_agent_front_end_on_the_gpu():
    """Row f4: BatchedRepresentationNetwork on CUDA, reading the DLPack observation block of a VecGame mid-episode, against
    the same module on the CPU fed with the oracle's observations (which tests/test_agent_frontend.py ties to the
    reference's own RepresentationNetwork).  Float32 convolutions: tolerance 1e-4 absolute / 1e-3 relative."""
    import torch
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    from xroute_env_b200.agent import BatchedRepresentationNetwork
    torch.manual_seed(3)
    geom = ispd18_geometry(30, 28, 9)
    insts = make_batch(geom, 4, 6, seed=77)
    vg = VecGame(geom, insts, device=0)
    vg.reset()
    orcs = [OracleEnv(geom, i) for i in insts]
    for net in (2, 5):
        vg.step(np.array([net] * 4, np.int32))
        for o in orcs:
            o.step(net)
    net_cpu = BatchedRepresentationNetwork().eval()
    net_gpu = BatchedRepresentationNetwork().eval()
    net_gpu.load_state_dict(net_cpu.state_dict())
    net_gpu = net_gpu.cuda()
    n_rem = vg.n_remaining.clone()
    with torch.no_grad():
        ob_g, rep_g, valid_g = net_gpu(vg.obs_batch()[:, :2 + 7 * 6], n_rem)
    obs = [o.obs()[0] for o in orcs]
    batch = np.zeros((4, 2 + 7 * 6) + obs[0].shape[1:], np.float32)
    for k, o in enumerate(obs):
        batch[k, : o.shape[0]] = o
    with torch.no_grad():
        ob_c, rep_c, valid_c = net_cpu(torch.from_numpy(batch), n_rem.cpu())
    assert torch.equal(valid_g.cpu(), valid_c) and int(valid_c.sum()) == 16
    assert torch.allclose(ob_g.cpu(), ob_c, atol=1e-4, rtol=1e-3)
    assert torch.allclose(rep_g.cpu(), rep_c, atol=1e-4, rtol=1e-3)
    vg.close()
