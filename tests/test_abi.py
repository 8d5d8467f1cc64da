"""The C-ABI library loads and exports exactly the symbols include/xroute_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "xroute_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xr_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from xroute_env_b200 import _lib
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from xroute_env_b200 import _lib
    L = _lib.load()
    for s in _header_symbols():
        assert hasattr(L, s), s
    assert L.xr_version() == 2


def test_config_struct_layout_matches_header():
    from xroute_env_b200._lib import XrConfig
    # 9 int32, 5 pointers (8-byte aligned), 5 + 1 + 7 int32
    assert ctypes.sizeof(XrConfig) == 136   # 36 (+4 pad) + 40 + 52 (+4 tail pad)
    assert XrConfig.x_coords.offset == 40 and XrConfig.via_cost.offset == 80


def test_config_struct_fields_agree_in_header_binding_and_integration_doc():
    """Field names and order of XrConfig: the header, the ctypes binding and the binding INTEGRATION.md prints."""
    src = open(os.path.join(ROOT, "include", "xroute_b200.h")).read()
    body = re.search(r"typedef struct XrConfig \{(.*?)\} XrConfig;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    hdr = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?\w+\s*\*?", "", decl, count=1)
        hdr += [n.strip().lstrip("*") for n in names.split(",")]
    from xroute_env_b200._lib import XrConfig
    assert hdr == [f[0] for f in XrConfig._fields_]
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blk = doc[doc.index("class XrConfig(C.Structure)"):]
    blk = blk[:blk.index("]\n") + 1]
    assert hdr == re.findall(r'\("(\w+)",', blk)


def test_no_oracle_reference_in_product_package():
    """The product path must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, "xroute_env_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "xr_oracle" not in txt, f


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    import numpy as np
    from xroute_env_b200._lib import XrError
    from xroute_env_b200.instances import ispd18_geometry, make_batch
    from xroute_env_b200.vec_game import VecGame
    g = ispd18_geometry(8, 8, 2)
    with pytest.raises((XrError, RuntimeError)):
        VecGame(g, make_batch(g, 1, 2, 0), device=0)
