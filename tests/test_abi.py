"""The C-ABI library loads on a CPU-only box and exports exactly the symbols include/mcb200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcb200.h")


def _header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from metacell_b200 import _lib
    L = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 38
    for name in declared:
        assert hasattr(L, name), f"{name} declared in mcb200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "metacell_b200/_lib.py SYMBOLS out of sync with the header"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (mcb_[a-z0-9_]+)", out)))
    assert exported == declared, set(exported) ^ set(declared)


def test_header_cites_reference_interfaces():
    txt = open(HEADER).read()
    for cite in ("R/utils.r:119", "R/utils.r:120", "R/utils.r:30-69", "R/scmat.r:399-407", "R/cgraph_knn.r:10-25",
                 "R/coclust_graph_cov.r:41"):
        assert cite in txt


def test_library_is_sm100a_with_tcgen05_and_tma():
    """SASS evidence that the correlation kernel is Blackwell-native (B200_PROFILING.md)"""
    from metacell_b200 import _lib
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    assert re.search(r"UTC[A-Z]*MMA", sass) and "UTMALDG" in sass and "LDTM" in sass


def test_no_gpu_means_loud_failure():
    from metacell_b200 import _lib
    L = _lib.lib()
    assert L.mcb_version() >= 100
    h = C.c_void_p()
    rc = L.mcb_ctx_create(0, C.byref(h))
    if rc == 0:
        L.mcb_ctx_destroy(h)
        pytest.skip("a GPU is present")
    msg = L.mcb_last_error().decode()
    assert "MC-ERR" in msg and "no CPU fallback" in msg
    with pytest.raises(_lib.McbError):
        _lib.Context(0)


def test_depth_rule_is_host_arithmetic(oracle):
    """mcb_downsamp_depth (reference R/scmat.r:399-407) needs no GPU and equals the oracle"""
    from metacell_b200 import _lib
    rng = np.random.default_rng(3)
    for n in (1, 2, 9, 1000, 12345):
        tot = rng.integers(100, 20000, size=n)
        assert _lib.downsamp_depth(tot) == oracle.downsamp_depth(tot)
    d = C.c_double(-1)
    assert _lib.lib().mcb_downsamp_depth(None, 0, C.byref(d)) == 0 and d.value == 0.0
    assert _lib.lib().mcb_downsamp_depth(None, 0, None) != 0            # null output is an error, not a crash


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "metacell_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mc_oracle" not in txt.replace("oracle/mc_oracle.c", ""), f
