"""Row f1: LEF/DEF/guide -> region instances.  Parser on a hand-written miniature design
(always runs), the committed ispd18_test1 fixtures against the design files when the
reference is mounted, and the CPU oracle over whole episodes of the real regions."""
import os

import numpy as np
import pytest

from helpers import GOLD, brute_force_dist
from xroute_env_b200.ispd import Design, extract_region, load_regions, parse_def, parse_guide, parse_lef, _orient

REF = "/root/reference/ispd/ispd18_test1/ispd18_test1.input"
FIX = os.path.join(GOLD, "ispd18_test1_regions.npz")

MINI_LEF = """
VERSION 5.8 ;
UNITS
  DATABASE MICRONS 1000 ;
END UNITS
LAYER M1
  TYPE ROUTING ;
  DIRECTION HORIZONTAL ;
  PITCH 0.2 0.2 ;
  WIDTH 0.06 ;
  SPACINGTABLE
    PARALLELRUNLENGTH 0
    WIDTH 0 0.06 ;
  SPACING 0.05 ;
END M1
LAYER V1
  TYPE CUT ;
  SPACING 0.07 ;
END V1
LAYER M2
  TYPE ROUTING ;
  DIRECTION VERTICAL ;
  PITCH 0.2 0.2 ;
  WIDTH 0.08 ;
  SPACING 0.06 ;
END M2
VIA V12 DEFAULT
  LAYER M1 ;
    RECT -0.05 -0.05 0.05 0.05 ;
  LAYER M2 ;
    RECT -0.05 -0.05 0.05 0.05 ;
END V12
MACRO INV
  SIZE 1.0 BY 2.0 ;
  PIN A
    DIRECTION INPUT ;
    USE SIGNAL ;
    PORT
      LAYER M1 ;
        RECT 0.15 0.35 0.25 0.45 ;
    END
  END A
  PIN Y
    DIRECTION OUTPUT ;
    PORT
      LAYER M1 ;
        RECT 0.75 1.55 0.85 1.65 ;
    END
  END Y
  PIN VDD
    USE POWER ;
    PORT
      LAYER M1 ;
        RECT 0.0 1.9 1.0 2.0 ;
    END
  END VDD
  OBS
    LAYER M2 ;
      RECT 0.4 0.9 0.6 1.1 ;
  END
END INV
END LIBRARY
"""
MINI_DEF = """
VERSION 5.8 ;
DESIGN mini ;
UNITS DISTANCE MICRONS 1000 ;
DIEAREA ( 0 0 ) ( 6000 4000 ) ;
TRACKS X 100 DO 30 STEP 200 LAYER M1 M2 ;
TRACKS Y 100 DO 20 STEP 200 LAYER M1 ;
TRACKS Y 150 DO 10 STEP 400 LAYER M2 ;
COMPONENTS 3 ;
- u1 INV + PLACED ( 1000 1000 ) N ;
- u2 INV + PLACED ( 3000 1000 ) FS ;
- u3 INV + PLACED ( 5000 0 ) N ;
END COMPONENTS
NETS 2 ;
- n1
  ( u1 Y ) ( u2 A )
 ;
- n2
  ( u2 Y ) ( u3 A )
 ;
END NETS
END DESIGN
"""
MINI_GUIDE = """n1
(
1000 2000 4000 3000 M1
)
n2
(
3000 0 6000 2000 M2
)
"""


@pytest.fixture()
def mini(tmp_path):
    for name, text in (("m.lef", MINI_LEF), ("m.def", MINI_DEF), ("m.guide", MINI_GUIDE)):
        (tmp_path / name).write_text(text)
    return Design.load(str(tmp_path / "m.lef"), str(tmp_path / "m.def"), str(tmp_path / "m.guide"))


def test_parsers_on_miniature_design(mini):
    lef, d = mini.lef, mini.deff
    assert [(l.name, l.direction, l.pitch, l.width, l.spacing) for l in lef.layers] == \
        [("M1", 0, 200, 60, 50), ("M2", 1, 200, 80, 60)]
    inv = lef.macros["INV"]
    assert inv.size == (1000, 2000) and set(inv.pins) == {"A", "Y", "VDD"}
    assert inv.pins["A"]["rects"] == [("M1", 150, 350, 250, 450)] and inv.pins["VDD"]["use"] == "POWER"
    assert inv.obs == [("M2", 400, 900, 600, 1100)]
    assert d.dbu == 1000 and d.die == (0, 0, 6000, 4000)
    assert d.tracks["M2"] == {"X": [(100, 30, 200)], "Y": [(150, 10, 400)]}
    assert d.components["u2"] == ("INV", 3000, 1000, "FS") and d.nets["n1"] == [("u1", "Y"), ("u2", "A")]
    assert mini.guides["n2"] == [(3000, 0, 6000, 2000, "M2")]


def test_orientation_transforms():
    r, size, at = (100, 200, 300, 500), (1000, 2000), (10, 20)
    assert _orient(r, size, at, "N") == (110, 220, 310, 520)
    assert _orient(r, size, at, "S") == (710, 1520, 910, 1820)
    assert _orient(r, size, at, "FN") == (710, 220, 910, 520)
    assert _orient(r, size, at, "FS") == (110, 1520, 310, 1820)
    assert _orient(r, size, at, "W") == (1510, 120, 1810, 320)
    assert _orient(r, size, at, "E") == (210, 720, 510, 920)
    assert _orient(r, size, at, "FW") == (210, 120, 510, 320)
    assert _orient(r, size, at, "FE") == (1510, 720, 1810, 920)


def test_region_extraction_miniature(mini):
    # box around u1/u2: n1 has both pins inside; n2 has one pin inside + a guide leaving to the east
    geom, inst = extract_region(mini, (800, 800, 4200, 3200), ext=200)
    assert geom.Z == 2 and list(geom.layer_dir) == [0, 1]
    assert geom.x_coords[0] == 700 and geom.x_coords[-1] == 4300 and np.all(np.diff(geom.x_coords) == 200)
    assert geom.y_coords[0] == 700 and geom.y_coords[-1] == 3300
    assert inst.net_ids == [1, 2]
    xs, ys = geom.x_coords, geom.y_coords
    aps = {(int(n), int(p)): (int(xs[x]), int(ys[y]), int(z)) for n, p, (x, y, z) in zip(inst.ap_net, inst.ap_pin, inst.ap_xyz)}
    # u1/Y at N (1000,1000): rect (1750..1850, 2550..2650) -> nearest crossing (1700|1900, 2500|2700); first AP listed wins the dict
    n1 = [(int(xs[x]), int(ys[y])) for n, (x, y, z) in zip(inst.ap_net, inst.ap_xyz) if n == 1]
    assert any(abs(px - 1800) <= 100 and abs(py - 2600) <= 100 for px, py in n1)
    # u2/A at FS (3000,1000): local (150..250, 350..450) -> y flipped: 1000 + 2000 - 450 = 2550..2650
    assert any(abs(px - 3200) <= 100 and abs(py - 2600) <= 100 for px, py in n1)
    # n2: boundary pin on the east edge, on M2 (z = 1)
    n2 = [(int(xs[x]), int(z)) for n, (x, y, z) in zip(inst.ap_net, inst.ap_xyz) if n == 2]
    assert (4100, 1) in n2
    # blockages: OBS of u1 on M2 and the power rails on M1; never on an AP
    blk = {tuple(int(v) for v in b) for b in inst.block_xyz}
    assert any(z == 1 for _, _, z in blk) and any(z == 0 for _, _, z in blk)
    assert not blk & {tuple(int(v) for v in a) for a in inst.ap_xyz}
    # without boundary pins n2 has a single pin in the box and is dropped
    _, inst2 = extract_region(mini, (800, 800, 4200, 3200), ext=200, boundary_pins=False)
    assert inst2.net_ids == [1]
    # union of all layers' tracks: M2's y tracks (150 + 400k) join -> non-uniform pitch
    g3, _ = extract_region(mini, (800, 800, 4200, 3200), ext=200, union_tracks=True)
    assert len(set(np.diff(g3.y_coords))) > 1


def test_fixture_invariants():
    regions = load_regions(FIX)
    assert {"t1_7x7_y79800", "t1_7x7_y319200", "t1_7x7_y79800_union"} <= set(regions)
    for name, (g, inst) in regions.items():
        assert g.Z == 9 and list(g.layer_dir) == [0, 1, 0, 1, 0, 1, 0, 1, 0], name
        assert np.all(np.diff(g.x_coords) > 0) and np.all(np.diff(g.y_coords) > 0)
        assert (inst.ap_xyz >= 0).all() and (inst.ap_xyz < [g.X, g.Y, g.Z]).all()
        cells = {tuple(c) for c in inst.ap_xyz.tolist()}
        assert len(cells) == len(inst.ap_xyz), "APs share a cell"
        assert not cells & {tuple(c) for c in inst.block_xyz.tolist()}
        for n in inst.net_ids:
            assert len(set(inst.ap_pin[inst.ap_net == n].tolist())) >= 2, (name, n)
    g, inst = regions["t1_7x7_y79800"]
    assert (g.X, g.Y) == (110, 115) and int(g.x_coords[0]) == 38200 and int(g.y_coords[0]) == 78090
    assert set(np.diff(g.x_coords)) == {400} and set(np.diff(g.y_coords)) == {380}
    gu, _ = regions["t1_7x7_y79800_union"]
    assert len(set(np.diff(gu.y_coords).tolist())) > 1      # M7-M9 tracks interleave: non-uniform pitch


@pytest.mark.skipif(not os.path.exists(REF + ".def"), reason="reference design files not mounted")
def test_fixture_matches_design_files():
    design = Design.load(REF + ".lef", REF + ".def", REF + ".guide")
    assert len(design.lef.macros) == 487 and len(design.deff.components) == 8879 and len(design.deff.nets) == 3153
    assert sum(len(v) for v in design.deff.nets.values()) == 17203
    assert sum(len(v) for v in design.guides.values()) == 26598
    regions = load_regions(FIX)
    for name, kw in (("t1_7x7_y79800", {}), ("t1_7x7_y319200", {}), ("t1_7x7_y79800_union", {"union_tracks": True})):
        g0, i0 = regions[name]
        g, inst = extract_region(design, i0.meta["route_box"], **kw)
        assert np.array_equal(g.x_coords, g0.x_coords) and np.array_equal(g.y_coords, g0.y_coords)
        assert np.array_equal(g.layer_pitch, g0.layer_pitch) and np.array_equal(g.layer_min_width, g0.layer_min_width)
        for f in ("block_xyz", "ap_net", "ap_pin", "ap_xyz"):
            assert np.array_equal(getattr(inst, f), getattr(i0, f)), (name, f)


@pytest.mark.parametrize("name", ["t1_7x7_y79800", "t1_7x7_y319200", "t1_7x7_y79800_union", "t1_1x1_gx3_gy6"])
def test_oracle_episode_on_real_regions(name):
    """configs[0]: one ispd18_test1 environment on the CPU, random net order: every net gets
    routed, metrics are cumulative and consistent with the occupancy, observation layout holds."""
    from oracle.oracle import OracleEnv
    g, inst = load_regions(FIX)[name]
    env = OracleEnv(g, inst)
    order = np.random.default_rng(7).permutation(inst.net_ids)
    wl = via = 0
    for k, net in enumerate(order):
        m = env.step(int(net))
        assert m["d_wirelength"] >= 0 and m["d_via"] >= 0
        wl += m["d_wirelength"]; via += m["d_via"]
        assert (m["wirelength"], m["via"]) == (wl, via) and m["violation"] == m["blocked"] + m["shorted"]
        cells, offs, costs = env.last_paths()
        assert len(costs) == len(set(inst.ap_pin[inst.ap_net == net].tolist())) - 1      # one connection per extra pin
        assert m["done"] == (k == len(order) - 1)
    usage, owner = env.state()
    assert m["overflow"] == int(np.maximum(usage.astype(np.int64) - 1, 0).sum())
    assert wl > 0 and env.obs().shape == (1, 2, g.Z, g.Y, g.X)


def test_oracle_distance_field_on_nonuniform_region():
    """A real 1x1 region laid on the union-of-tracks coordinates (non-uniform y pitch), two nets
    already routed: the oracle's field equals the independent numpy Bellman-Ford."""
    from oracle.oracle import OracleEnv
    from xroute_env_b200.instances import Geometry
    regions = load_regions(FIX)
    g, inst = regions["t1_1x1_gx3_gy6"]
    gu = regions["t1_7x7_y79800_union"][0]
    y0 = int(np.searchsorted(gu.y_coords, g.y_coords[0]))
    gn = Geometry(X=g.X, Y=g.Y, Z=g.Z, x_coords=g.x_coords, y_coords=gu.y_coords[y0:y0 + g.Y].copy(),
                  layer_dir=g.layer_dir, layer_pitch=g.layer_pitch, layer_min_width=g.layer_min_width)
    assert len(set(np.diff(gn.y_coords).tolist())) > 1
    env = OracleEnv(gn, inst)
    env.step(inst.net_ids[1]); env.step(inst.net_ids[2])
    net = inst.net_ids[0]
    usage, _ = env.state()
    apnet = np.zeros((g.Z, g.Y, g.X), np.int64)
    apnet[inst.ap_xyz[:, 2], inst.ap_xyz[:, 1], inst.ap_xyz[:, 0]] = inst.ap_net
    blk = np.zeros((g.Z, g.Y, g.X), np.uint8)
    blk[inst.block_xyz[:, 2], inst.block_xyz[:, 1], inst.block_xyz[:, 0]] = 1
    cflag = (usage > 0).astype(np.uint8) | (((apnet != 0) & (apnet != net)).astype(np.uint8) << 1) | (blk << 2)
    srcs = [tuple(int(v) for v in c) for c, n, p in zip(inst.ap_xyz, inst.ap_net, inst.ap_pin) if n == net and p == 1]
    got = env.distance_field(net, [(z * g.Y + y) * g.X + x for x, y, z in srcs]).astype(np.int64)
    assert np.array_equal(got, brute_force_dist(gn, cflag, srcs))
