#!/usr/bin/env python
"""Regenerate tests/golden/ispd18_test1_regions.npz from the reference's benchmark files.

    python tools/make_ispd_regions.py [/root/reference/ispd/ispd18_test1/ispd18_test1.input]

Regions: the two 7x7-gcell worker boxes the reference ships dumps for
(/root/reference/ispd/ispd18_test1/dump/workerx39900_y79800, .../workerx39900_y319200; boxes from
the worker.bin headers), the first one also on the union-of-all-layers
track grid (non-uniform pitch), and the four busiest 1x1 gcells inside the first box.
The design files do not travel to the GPU box; the extracted plain arrays do.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xroute_env_b200.ispd import Design, extract_region, save_regions  # noqa: E402

GCELL = 5700
BOX_A = (39900, 79800, 79800, 119700)
BOX_B = (39900, 319200, 79800, 359100)


def main():
    stem = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/ispd/ispd18_test1/ispd18_test1.input"
    design = Design.load(stem + ".lef", stem + ".def", stem + ".guide")
    regions = {
        "t1_7x7_y79800": extract_region(design, BOX_A),
        "t1_7x7_y319200": extract_region(design, BOX_B),
        "t1_7x7_y79800_union": extract_region(design, BOX_A, union_tracks=True),
    }
    cells = []
    for gy in range(7):
        for gx in range(7):
            box = (BOX_A[0] + gx * GCELL, BOX_A[1] + gy * GCELL, BOX_A[0] + (gx + 1) * GCELL, BOX_A[1] + (gy + 1) * GCELL)
            try:
                g, inst = extract_region(design, box)
            except ValueError:
                continue
            cells.append((-len(inst.net_ids), gx, gy, g, inst))
    cells.sort(key=lambda c: c[:3])
    for n, gx, gy, g, inst in cells[:4]:
        regions[f"t1_1x1_gx{gx}_gy{gy}"] = (g, inst)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ispd18_test1_regions.npz")
    save_regions(out, regions)
    for name, (g, inst) in regions.items():
        print(f"{name}: {g.X}x{g.Y}x{g.Z}, {len(inst.net_ids)} nets, {len(inst.ap_net)} APs, {len(inst.block_xyz)} blockages")
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
