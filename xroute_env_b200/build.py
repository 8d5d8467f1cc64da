"""Build libxroute_b200.so in-tree with nvcc for sm_100a (python -m xroute_env_b200.build).

Two translation units (the C ABI + sweep engines, and the frontier engine) are compiled side by side into
csrc/_obj/*.o and linked into one shared library; a unit is rebuilt only when one of its sources is newer."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
UNITS = ["xr_api.cu", "xr_frontier.cu"]
OUT = os.path.join(HERE, "libxroute_b200.so")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "xroute_b200.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("XR_NVCC_EXTRA", "").split()      # e.g. -DFR_TIMING for tools/diag_frontier.py
    os.makedirs(OBJ, exist_ok=True)
    newest = max(os.path.getmtime(d) for d in _deps())
    tag = os.path.join(OBJ, "flags.txt")
    flags = " ".join(extra)
    if not os.path.exists(tag) or open(tag).read() != flags:
        force = True
    procs, objs = [], []
    for u in UNITS:
        o = os.path.join(OBJ, u.replace(".cu", ".o"))
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) >= newest:
            continue
        cmd = [nvcc] + extra + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                                "-Xcompiler", "-fPIC", "-c", "-o", o, os.path.join(CSRC, u)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((u, subprocess.Popen(cmd)))
    failed = [u for u, p in procs if p.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, f"nvcc {failed}")
    if procs or force or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs)
    with open(tag, "w") as f:
        f.write(flags)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
