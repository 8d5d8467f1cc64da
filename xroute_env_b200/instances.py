"""Region instances: grid geometry presets and the seeded synthetic generator.

A *region instance* is what the reference's simulator would load from a
``dump/worker*`` directory (``/root/reference/examples/launch_training.py:33-62``)
and describe to the agent as a node stream (``net_ordering.proto:11-45``): the
track grid, its blockages, and the access points (APs) of every pin of every net
that has to be routed inside the region.

Geometry follows ispd18_test1 (``ispd/ispd18_test1/ispd18_test1.input.def:234-251``
tracks, ``ispd18_test1.input.lef:13-185`` layers): x tracks every 400 DBU from
200, y tracks every 380 DBU from 190, nine routing layers alternating
HORIZONTAL/VERTICAL from Metal1, pitch 380/400/.../660 DBU, width 120/140 DBU.

The generator is host-side numpy only; it feeds *both* the CUDA library and the
CPU oracle with the same plain arrays, so neither depends on the other.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Router cost constants pinned by ispd/ispd18_test1/dump/init_globals.bin and
# dump/workerx39900_y79800/worker.bin.
VIACOST = 4
GRIDCOST = 2
DRCCOST = 8          # workerDRCCost
FIXEDSHAPECOST = 8   # workerFixedShapeCost
BLOCKCOST = 32

# Grid presets.
PRESETS = {
    "T1-1x1": (25, 26, 9),
    "T1-7x7": (112, 116, 9),
    "SYN-256": (256, 256, 9),
    "SYN-1024": (1024, 1024, 9),
}

# ispd18_test1 net-degree histogram (2:1950 3:104 4:672 5:63 6:28, tail to 66),
# tail bucketed.
_DEGREES = np.array([2, 3, 4, 5, 6, 8, 12, 17, 30])
_DEGREE_P = np.array([0.618, 0.033, 0.213, 0.020, 0.009, 0.050, 0.030, 0.020, 0.007])
_DEGREE_P = _DEGREE_P / _DEGREE_P.sum()


@dataclass
class Geometry:
    """Track grid shared by every environment of a batch."""
    X: int
    Y: int
    Z: int
    x_coords: np.ndarray          # int32 [X]  DBU
    y_coords: np.ndarray          # int32 [Y]
    layer_dir: np.ndarray         # uint8 [Z]  0 = horizontal (preferred axis x)
    layer_pitch: np.ndarray       # int32 [Z]
    layer_min_width: np.ndarray   # int32 [Z]
    via_cost: int = VIACOST
    grid_cost: int = GRIDCOST
    drc_cost: int = DRCCOST
    fixed_shape_cost: int = FIXEDSHAPECOST
    block_cost: int = BLOCKCOST

    @property
    def cells(self) -> int:
        return self.X * self.Y * self.Z


def ispd18_geometry(X: int, Y: int, Z: int = 9) -> Geometry:
    pitch = np.full(Z, 400, np.int32)
    width = np.full(Z, 140, np.int32)
    pitch[0], width[0] = 380, 120
    if Z >= 9:
        pitch[8] = 660
    return Geometry(
        X=X, Y=Y, Z=Z,
        x_coords=(200 + 400 * np.arange(X)).astype(np.int32),
        y_coords=(190 + 380 * np.arange(Y)).astype(np.int32),
        layer_dir=(np.arange(Z) % 2).astype(np.uint8),
        layer_pitch=pitch, layer_min_width=width,
    )


def preset_geometry(name: str) -> Geometry:
    return ispd18_geometry(*PRESETS[name])


@dataclass
class Instance:
    """One region: blockages + access points (parallel arrays)."""
    block_xyz: np.ndarray                 # int32 [n_block, 3]
    ap_net: np.ndarray                    # int32 [n_ap]  1-based net id
    ap_pin: np.ndarray                    # int32 [n_ap]  1-based pin id inside the net
    ap_xyz: np.ndarray                    # int32 [n_ap, 3]
    meta: dict = field(default_factory=dict)
    guides: np.ndarray | None = None      # int32 [n, 6] net, x0, x1, y0, y1, z (cells, inclusive): route guides (optional)

    @property
    def net_ids(self) -> list[int]:
        return sorted(set(int(v) for v in self.ap_net))


def make_instance(geom: Geometry, n_nets: int, seed: int, *, p_obstacle: float = 0.10,
                  hot_spots: int = 0, hot_sigma: float = 32.0,
                  max_degree: int | None = None) -> Instance:
    """Seeded synthetic region.

    ``rng = PCG64(seed)``; blockages Bernoulli(p) on the two lowest layers and
    p/5 above; net degree drawn from the ispd18_test1 histogram; every pin has
    1-3 APs on x-adjacent cells at z in {0, 1}; APs never share a cell and never
    sit on a blockage.  ``hot_spots > 0`` clusters the net centres (Gaussian,
    ``hot_sigma`` cells) for the congestion-heavy configuration.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    X, Y, Z = geom.X, geom.Y, geom.Z
    pz = np.where(np.arange(Z) <= 1, p_obstacle, p_obstacle / 5.0)
    block = rng.random((Z, Y, X)) < pz[:, None, None]
    taken = np.zeros((Z, Y, X), bool)
    span = max(4, min(X, Y) // 4)
    if hot_spots > 0:
        hx = rng.integers(0, X, hot_spots)
        hy = rng.integers(0, Y, hot_spots)
    degs = _DEGREES if max_degree is None else _DEGREES[_DEGREES <= max_degree]
    degp = _DEGREE_P[: len(degs)] / _DEGREE_P[: len(degs)].sum()
    ap_net, ap_pin, ap_xyz = [], [], []
    for net in range(1, n_nets + 1):
        deg = int(rng.choice(degs, p=degp))
        if hot_spots > 0:
            h = int(rng.integers(0, hot_spots))
            cx = int(np.clip(round(hx[h] + rng.normal(0, hot_sigma)), 0, X - 1))
            cy = int(np.clip(round(hy[h] + rng.normal(0, hot_sigma)), 0, Y - 1))
        else:
            cx, cy = int(rng.integers(0, X)), int(rng.integers(0, Y))
        for pin in range(1, deg + 1):
            for _attempt in range(64):
                x = int(np.clip(cx + rng.integers(-span, span + 1), 0, X - 1))
                y = int(np.clip(cy + rng.integers(-span, span + 1), 0, Y - 1))
                z = int(rng.integers(0, min(2, Z)))
                n_ap = int(rng.integers(1, 4))
                xs = [xx for xx in range(x, min(X, x + n_ap)) if not taken[z, y, xx]]
                if xs and xs[0] == x:
                    # keep the run contiguous from x
                    run = []
                    for xx in xs:
                        if run and xx != run[-1] + 1:
                            break
                        run.append(xx)
                    for xx in run:
                        taken[z, y, xx] = True
                        block[z, y, xx] = False
                        ap_net.append(net); ap_pin.append(pin); ap_xyz.append((xx, y, z))
                    break
            else:
                raise RuntimeError("could not place pin (grid too full)")
    bz, by, bx = np.nonzero(block)
    return Instance(
        block_xyz=np.stack([bx, by, bz], 1).astype(np.int32).reshape(-1, 3),
        ap_net=np.asarray(ap_net, np.int32), ap_pin=np.asarray(ap_pin, np.int32),
        ap_xyz=np.asarray(ap_xyz, np.int32).reshape(-1, 3),
        meta={"seed": seed, "n_nets": n_nets, "p_obstacle": p_obstacle, "hot_spots": hot_spots},
    )


def make_batch(geom: Geometry, n_envs: int, n_nets: int, seed: int, *, first_env: int = 0,
               **kw) -> list[Instance]:
    """``n_envs`` instances; env i uses ``PCG64(seed + first_env + i)`` so a shard of
    a larger batch (one rank of a multi-GPU job) generates exactly its slice."""
    return [make_instance(geom, n_nets, seed + first_env + i, **kw) for i in range(n_envs)]


def export_data(geom: Geometry, inst: Instance, usage: np.ndarray, cum_metrics=(0, 0, 0),
                net_list=None) -> list:
    """The ``data`` list the reference's ``handle_messange`` produces
    (``/root/reference/baseline/baseline_utils.py:16-40``) for this region in the
    occupancy state ``usage`` (uint8 [Z, Y, X], canonical layout):
    ``[[X,Y,Z], nodes, [vio, wl, via], nets]`` with node
    ``[[mx,my,mz],[px,py,pz],[used, Net, Pin]]``; ``Net`` -1 blockage, 0 normal,
    >=1 access point.  Used to drive the reference's own ``build_3Dgrid`` in the
    differential tests and golden-vector generation."""
    X, Y, Z = geom.X, geom.Y, geom.Z
    net = np.zeros((Z, Y, X), np.int64)
    pin = np.full((Z, Y, X), -1, np.int64)
    for n, p, (x, y, z) in zip(inst.ap_net, inst.ap_pin, inst.ap_xyz):
        net[z, y, x] = n
        pin[z, y, x] = p
    for x, y, z in inst.block_xyz:
        net[z, y, x] = -1
    used = (usage.reshape(Z, Y, X) > 0) | (net == -1)
    nodes = []
    for z in range(Z):
        for y in range(Y):
            for x in range(X):
                nodes.append([[x, y, z], [int(geom.x_coords[x]), int(geom.y_coords[y]), z],
                              [int(used[z, y, x]), int(net[z, y, x]), int(pin[z, y, x])]])
    if net_list is None:
        net_list = inst.net_ids
    return [[X, Y, Z], nodes, [int(v) for v in cum_metrics], list(net_list)]
