"""Build libxroute_b200.so in-tree with nvcc for sm_100a (python -m xroute_env_b200.build)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "xr_api.cu")
OUT = os.path.join(HERE, "libxroute_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + \
       [os.path.join(HERE, "..", "include", "xroute_b200.h")]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    cmd[1:1] = os.environ.get("XR_NVCC_EXTRA", "").split()      # e.g. -DWIN_PHASE_TIMING for tools/diag_route.py
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
