"""The maze part of the oracle has no reference vectors (PARITY UNPINNED, see
oracle/xr_oracle.c header); it is checked here against (a) hand-made mazes with known
answers and (b) an independent dense numpy Bellman-Ford restatement of the SPEC."""
import numpy as np
import pytest

from helpers import brute_force_dist, spec_route_net
from oracle.oracle import OracleEnv
from xroute_env_b200.instances import Instance, ispd18_geometry, make_instance


def _inst(block, aps):
    return Instance(block_xyz=np.array(block, np.int32).reshape(-1, 3),
                    ap_net=np.array([a[0] for a in aps], np.int32), ap_pin=np.array([a[1] for a in aps], np.int32),
                    ap_xyz=np.array([a[2] for a in aps], np.int32).reshape(-1, 3))


def ci(g, x, y, z):
    return (z * g.Y + y) * g.X + x


def test_straight_preferred_direction():
    g = ispd18_geometry(10, 6, 3)                       # layer 0 horizontal
    env = OracleEnv(g, _inst([], [(1, 1, (1, 2, 0)), (1, 2, (7, 2, 0))]))
    m = env.step(1)
    cells, off, cost = env.last_paths()
    assert cost.tolist() == [6 * 400] and m["d_wirelength"] == 2400 and m["d_via"] == 0 and m["d_violation"] == 0
    assert len(cells) == 7 and all(c == ci(g, x, 2, 0) for c, x in zip(cells, range(7, 0, -1))) or \
        all(c == ci(g, x, 2, 0) for c, x in zip(cells, range(1, 8)))


def test_wrong_way_uses_upper_layer_when_cheaper():
    g = ispd18_geometry(6, 12, 3)                       # y move on layer 0 costs 3x; layer 1 is vertical
    env = OracleEnv(g, _inst([], [(1, 1, (2, 1, 0)), (1, 2, (2, 10, 0))]))
    m = env.step(1)
    _, _, cost = env.last_paths()
    wrong_way = 9 * 380 * 3
    via_route = 2 * 4 * 400 + 9 * 380
    assert cost[0] == min(wrong_way, via_route) == via_route
    assert m["d_via"] == 2 and m["d_wirelength"] == 9 * 380


def test_blockage_is_detoured_not_crossed():
    g = ispd18_geometry(9, 5, 1)
    block = [(4, y, 0) for y in range(0, 4)]            # wall with a gap at y = 4
    env = OracleEnv(g, _inst(block, [(1, 1, (1, 1, 0)), (1, 2, (7, 1, 0))]))
    m = env.step(1)
    cells, _, cost = env.last_paths()
    assert m["blocked"] == 0 and m["d_violation"] == 0
    detour = 6 * 400 + 2 * 3 * 380 * 3                  # 6 x-steps + 3 up + 3 down wrong-way on a horizontal layer
    assert cost[0] == detour
    assert ci(g, 4, 4, 0) in set(cells.tolist())


def test_short_through_other_net_counts_violation():
    g = ispd18_geometry(7, 3, 1)
    # net 1 builds a vertical wall x=3 (wrong-way, only layer); net 2 must cross it
    aps = [(1, 1, (3, 0, 0)), (1, 2, (3, 2, 0)), (2, 1, (0, 1, 0)), (2, 2, (6, 1, 0))]
    env = OracleEnv(g, _inst([], aps))
    m1 = env.step(1)
    assert m1["d_violation"] == 0
    m2 = env.step(2)
    assert m2["shorted"] == 1 and m2["overflow"] == 1 and m2["d_violation"] == 1 and m2["done"] == 1
    _, _, cost = env.last_paths()
    assert cost[0] == 5 * 400 + 400 * (1 + 8)            # one cell at drc cost 8


def test_multi_pin_tree_and_source_pin():
    g = ispd18_geometry(11, 11, 2)
    aps = [(1, 1, (0, 5, 0)), (1, 2, (5, 5, 0)), (1, 3, (10, 5, 0)), (1, 4, (5, 0, 1))]
    env = OracleEnv(g, _inst([], aps))
    assert env.src_pin(1) == 2                          # the pin at the bbox centre
    m = env.step(1)
    cells, off, cost = env.last_paths()
    assert len(cost) == 3 and m["done"] == 1
    assert sorted(cost.tolist()) == sorted([2000, 2000, 4 * 400 + 5 * 380])
    usage, owner = env.state()
    assert usage.sum() == len(set(cells.tolist())) and (owner[usage > 0] == 1).all()


def test_tie_break_is_canonical():
    g = ispd18_geometry(5, 5, 1)
    # two targets at equal distance: the smaller cell index wins
    aps = [(1, 1, (2, 2, 0)), (1, 2, (0, 2, 0)), (1, 3, (4, 2, 0))]
    env = OracleEnv(g, _inst([], aps))
    env.step(1)
    cells, off, cost = env.last_paths()
    assert cost.tolist() == [800, 800]
    assert cells[off[0]] == ci(g, 0, 2, 0) and cells[off[1]] == ci(g, 4, 2, 0)


def test_backtrace_reads_the_converged_field_only():
    """x pitch 300 / y pitch 100 on a horizontal layer: an x step costs what a wrong-way y step costs (100 * 3).  The walk
    back from (2,6) runs along row 6 to (5,6); there the cell it just left is one x step away, and so is the source
    (5,5) one y step below.  The path is read from the converged distance field, so it never turns back: it ends on
    the source (the GPU engines once differed here, tests/test_gpu_parity.py::test_backtrace_ignores_...)."""
    g = ispd18_geometry(14, 16, 3)
    g.x_coords = (300 * np.arange(14)).astype(np.int32)
    g.y_coords = (100 * np.arange(16)).astype(np.int32)
    aps = [(1, 1, (5, 5, 0)), (1, 2, (2, 6, 0)), (2, 1, (9, 10, 0)), (2, 2, (12, 10, 0)), (2, 3, (7, 11, 0))]
    env = OracleEnv(g, _inst([], aps))
    env.step(1)
    cells, off, cost = env.last_paths()
    assert cost.tolist() == [1200]
    assert cells.tolist() == [ci(g, 2, 6, 0), ci(g, 3, 6, 0), ci(g, 4, 6, 0), ci(g, 5, 6, 0), ci(g, 5, 5, 0)]
    env.step(2)
    cells, off, cost = env.last_paths()
    assert cost.tolist() == [900, 900]
    assert cells[off[1]:].tolist() == [ci(g, 7, 11, 0), ci(g, 8, 11, 0), ci(g, 9, 11, 0), ci(g, 9, 10, 0)]


@pytest.mark.parametrize("seed", range(6))
def test_distance_field_matches_brute_force(seed):
    rng = np.random.default_rng(seed)
    X, Y, Z = int(rng.integers(4, 14)), int(rng.integers(4, 14)), int(rng.integers(1, 6))
    g = ispd18_geometry(X, Y, Z)
    if seed % 2:
        g.x_coords = np.cumsum(rng.integers(50, 900, X)).astype(np.int32)
        g.y_coords = np.cumsum(rng.integers(50, 900, Y)).astype(np.int32)
    inst = make_instance(g, 5, seed, p_obstacle=0.3)
    env = OracleEnv(g, inst)
    order = [int(v) for v in rng.permutation(inst.net_ids)]
    for net in order[:3]:
        env.step(net)
    net = order[3]
    usage, _ = env.state()
    apnet = np.zeros((Z, Y, X), np.int64)
    for n, (x, y, z) in zip(inst.ap_net, inst.ap_xyz):
        apnet[z, y, x] = n
    blk = np.zeros((Z, Y, X), np.uint8)
    for x, y, z in inst.block_xyz:
        blk[z, y, x] = 1
    cflag = ((usage > 0).astype(np.uint8) | (((apnet != 0) & (apnet != net)).astype(np.uint8) << 1) | (blk << 2))
    srcs = [tuple(int(v) for v in inst.ap_xyz[i]) for i in range(len(inst.ap_net)) if inst.ap_net[i] == net][:2]
    want = brute_force_dist(g, cflag, srcs)
    got = env.distance_field(net, [ci(g, *s) for s in srcs]).astype(np.int64)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", range(4))
def test_paths_are_consistent_with_costs_and_metrics(seed):
    """Every committed path is connected, its weight sum equals the reported connection cost,
    early-terminated and full Dijkstra agree, and the deltas add up to the cumulative metrics."""
    g = ispd18_geometry(20, 18, 4)
    inst = make_instance(g, 8, 40 + seed, p_obstacle=0.25)
    a, b = OracleEnv(g, inst), OracleEnv(g, inst)
    tot = np.zeros(3, np.int64)
    for net in np.random.default_rng(seed).permutation(inst.net_ids):
        ma, mb = a.step(int(net)), b.step(int(net), full=True)
        assert ma == mb
        pa, pb = a.last_paths(), b.last_paths()
        assert all(np.array_equal(x, y) for x, y in zip(pa, pb))
        cells, off, cost = pa
        wl = via = 0
        for k in range(len(cost)):
            seg = cells[off[k]:off[k + 1]]
            for u, v in zip(seg[:-1], seg[1:]):
                ux, uy, uz = u % g.X, (u // g.X) % g.Y, u // (g.X * g.Y)
                vx, vy, vz = v % g.X, (v // g.X) % g.Y, v // (g.X * g.Y)
                assert abs(ux - vx) + abs(uy - vy) + abs(uz - vz) == 1
                via += uz != vz
                wl += abs(int(g.x_coords[ux]) - int(g.x_coords[vx])) + abs(int(g.y_coords[uy]) - int(g.y_coords[vy]))
        assert (ma["d_wirelength"], ma["d_via"]) == (wl, via)
        tot += [ma["d_violation"], ma["d_wirelength"], ma["d_via"]]
        assert tot.tolist() == [ma["violation"], ma["wirelength"], ma["via"]]
        assert ma["violation"] == ma["blocked"] + ma["shorted"]
    with pytest.raises(ValueError):
        a.step(int(inst.net_ids[0]))                    # already routed -> illegal


@pytest.mark.parametrize("seed", range(8))
def test_whole_net_routing_matches_independent_restatement(seed):
    """Targets, canonical backtrace, commit and the tree bookkeeping against a second restatement of the SPEC that shares
    no code with the C oracle (dense Bellman-Ford + plain-Python walk), on grids made of ties: every pitch and via
    cost a small multiple of 100, random layer directions and cost constants."""
    rng = np.random.default_rng(100 + seed)
    X, Y, Z = int(rng.integers(5, 13)), int(rng.integers(5, 13)), int(rng.integers(2, 5))
    g = ispd18_geometry(X, Y, Z)
    px, py = rng.choice([100, 200, 300], 2)
    g.x_coords = (np.cumsum(rng.choice([1, 1, 2, 3], X)) * px).astype(np.int32)
    g.y_coords = (py * np.arange(Y)).astype(np.int32)
    g.layer_dir = rng.integers(0, 2, Z).astype(np.uint8)
    g.layer_pitch = rng.choice([25, 50, 75, 100], Z).astype(np.int32)
    g.layer_min_width = rng.choice([5, 10, 20], Z).astype(np.int32)
    g.via_cost, g.grid_cost = int(rng.choice([1, 2, 4])), int(rng.choice([0, 1, 2]))
    g.drc_cost, g.fixed_shape_cost, g.block_cost = int(rng.choice([1, 2, 8])), int(rng.choice([1, 2, 8])), int(rng.choice([1, 5, 32]))
    inst = make_instance(g, 4, 300 + seed, p_obstacle=0.2)
    env = OracleEnv(g, inst)
    usage, owner = np.zeros((Z, Y, X), np.uint8), np.zeros((Z, Y, X), np.uint16)
    for net in rng.permutation(inst.net_ids):
        m = env.step(int(net))
        cells, off, cost, wl, via = spec_route_net(g, inst, usage, owner, int(net))
        oc, oo, ocost = env.last_paths()
        assert ocost.tolist() == cost and oo.tolist() == off and oc.tolist() == cells, (seed, net)
        assert (m["d_wirelength"], m["d_via"]) == (wl, via)
        ou, oown = env.state()
        assert np.array_equal(ou, usage) and np.array_equal(oown, owner)


@pytest.mark.parametrize("seed", range(10))
def test_goal_directed_partial_field_gives_the_same_routes(seed):
    """Ground work for a goal-directed GPU search: a best-first search on d + h that stops at the
    best target distance B leaves only the cells with d + h <= B final (1-3 % of a window on the bench workload,
    tools/analyze_pruning.py) -- target choice, canonical walk, commit and metrics must not change.  Tie-heavy grids,
    obstacles and foreign wires included; counted: the share of cells the partial search left unsettled."""
    rng = np.random.default_rng(700 + seed)
    X, Y, Z = int(rng.integers(8, 22)), int(rng.integers(8, 22)), int(rng.integers(2, 5))
    g = ispd18_geometry(X, Y, Z)
    if seed % 2:
        px, py = rng.choice([100, 200, 300], 2)
        g.x_coords = (np.cumsum(rng.choice([1, 1, 2, 3], X)) * px).astype(np.int32)
        g.y_coords = (py * np.arange(Y)).astype(np.int32)
        g.layer_dir = rng.integers(0, 2, Z).astype(np.uint8)
        g.layer_pitch = rng.choice([25, 50, 75, 100], Z).astype(np.int32)
        g.via_cost, g.grid_cost = int(rng.choice([1, 2, 4])), int(rng.choice([0, 1, 2]))
    inst = make_instance(g, 6, 800 + seed, p_obstacle=0.2)
    env = OracleEnv(g, inst)
    usage, owner = np.zeros((Z, Y, X), np.uint8), np.zeros((Z, Y, X), np.uint16)
    for net in rng.permutation(inst.net_ids):
        m = env.step(int(net))
        cells, off, cost, wl, via = spec_route_net(g, inst, usage, owner, int(net), goal_directed=True)
        oc, oo, ocost = env.last_paths()
        assert ocost.tolist() == cost and oo.tolist() == off and oc.tolist() == cells, (seed, net)
        assert (m["d_wirelength"], m["d_via"]) == (wl, via)
    ou, oown = env.state()
    assert np.array_equal(ou, usage) and np.array_equal(oown, owner)


def _one_layer(X, Y, aps, guides=None):
    from xroute_env_b200.instances import Instance
    geom = ispd18_geometry(X, Y, 1)                      # one horizontal layer: x step 400, y step 380 (wrong-way: x3)
    a = np.array(aps, np.int32)
    inst = Instance(block_xyz=np.zeros((0, 3), np.int32), ap_net=a[:, 0].copy(), ap_pin=a[:, 1].copy(),
                    ap_xyz=np.ascontiguousarray(a[:, 2:5]))
    if guides is not None:
        inst.guides = np.array(guides, np.int32).reshape(-1, 6)
    return geom, inst


def test_guide_term_known_answers():
    """Hand-made maze for the out-of-guide term (run-net-ordering-training.tcl:3 -follow_guide 1, GUIDECOST): pins at both
    ends of row 0, the guide runs around through row 2.  Straight: 2 + 3 steps inside the guide (400 each) and 3 outside
    (400 (1 + GUIDE)); around: 4 x steps inside + 4 inside x steps of the end pieces (400 each) and 4 wrong-way y steps
    (380 x 3).  GUIDE = 1: straight wins, 8 x 400 + 3 x 400 = 4400.  GUIDE = 20: around wins, 8 x 400 + 4 x 1140 = 7760."""
    guides = [(1, 0, 2, 0, 0, 0), (1, 2, 2, 0, 2, 0), (1, 2, 6, 2, 2, 0), (1, 6, 6, 0, 2, 0), (1, 6, 8, 0, 0, 0)]
    geom, inst = _one_layer(9, 5, [(1, 1, 0, 0, 0), (1, 2, 8, 0, 0)], guides)
    for gcost, want, rows in ((0, 3200, {0}), (1, 4400, {0}), (20, 7760, {0, 1, 2})):
        env = OracleEnv(geom, inst, guide_cost=gcost)
        env.step(1)
        cells, off, cost = env.last_paths()
        assert cost.tolist() == [want], (gcost, cost)
        assert {int(c) // 9 for c in cells} == rows, (gcost, cells)


def test_halo_term_known_answers():
    """Hand-made maze for the spacing halo (SHAPEBLOATWIDTH): net 1 runs straight along row 3; net 2 has its pins at both
    ends of row 2, the track next to it.  Without halo net 2 runs straight (9 x 400 = 3600).  With halo 1 every cell of
    row 2 counts as route shape (x 9): net 2 drops to row 1 (wrong-way 380 x 3), runs there (9 x 400) and climbs back
    into the halo at its pin (380 x (1 + 2 + 8)): 1140 + 3600 + 4180 = 8920.  The metrics do not see the halo."""
    geom, inst = _one_layer(10, 7, [(1, 1, 0, 3, 0), (1, 2, 9, 3, 0), (2, 1, 0, 2, 0), (2, 2, 9, 2, 0)])
    for halo, want, rows in ((0, 3600, {2}), (1, 8920, {1, 2})):
        env = OracleEnv(geom, inst, halo=halo)
        m1 = env.step(1)
        assert env.last_paths()[2].tolist() == [3600]
        m2 = env.step(2)
        cells, off, cost = env.last_paths()
        assert cost.tolist() == [want], (halo, cost)
        assert {int(c) // 10 for c in cells} == rows
        assert (m2["violation"], m2["overflow"]) == (0, 0)
