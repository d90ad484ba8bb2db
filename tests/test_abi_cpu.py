"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "drake_b200_hydro.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dbh_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from drake_b200 import build, engine
    path = build.build()
    assert os.path.exists(path)
    L = engine.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/drake_b200_hydro.h but not exported"
    assert L.dbh_abi_version() == 1000


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import drake_b200 as db
    with pytest.raises(db.HydroError) as ei:
        db.Context(0)
    assert ei.value.status == db.engine.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_the_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "drake_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("no cpu or oracle", ""), os.path.join(dirpath, f)


def test_struct_layouts_match_header():
    from drake_b200 import engine
    assert ctypes.sizeof(engine.AssetInfo) == 8 * 4 + 6 * 8
    assert ctypes.sizeof(engine.BatchInfo) == 10 * 4
    assert ctypes.sizeof(engine.ResultSizes) == 5 * 8 + 2 * 4
    assert ctypes.sizeof(engine.ResultBuffers) == 20 * 8
    assert ctypes.sizeof(engine.Stats) == 10 * 8 + 14 * 8
