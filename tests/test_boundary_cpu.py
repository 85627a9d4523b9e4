"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/cnnacc.h declares, and fails loudly (no fallback) when there is no GPU.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

import fpga_cnn_b200 as fc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    fc.build()
    return fc.load()


def test_header_and_library_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "cnnacc.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\s*\*\s*|int64_t\s+|int\s+)(cnnacc_\w+|cnn_infer)\s*\(", hdr, re.M))
    assert declared, "no declarations parsed"
    assert declared == set(fc._lib.SYMBOLS), declared ^ set(fc._lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_reference_symbol_convention(lib):
    """ARMEngine loads cnn_infer with argtypes=[c_void_p]*4, restype=c_int (realtime_detect.py:389-391)."""
    arm = fc.load_arm_cnn_lib()
    assert arm.cnn_infer.argtypes == [ctypes.c_void_p] * 4 and arm.cnn_infer.restype is ctypes.c_int


def test_null_handle_calls_are_rejected(lib):
    assert lib.cnnacc_set_shifts(None, 2, 4, 6) == fc._lib.ERR_ARG
    assert lib.cnnacc_destroy(None) == fc._lib.ERR_ARG
    assert lib.cnnacc_launch_count(None) == 0
    assert lib.cnn_infer(None, None, None, None) == fc._lib.ERR_ARG


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        fc.CNNAccelerator()
    # the drop-in symbol must also refuse rather than compute on the CPU
    img = np.zeros(16384, np.uint8); wt = np.zeros(23184, np.uint8); sh = np.array([2, 4, 6], np.int32); out = np.zeros(16384, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.cnn_infer(p(img), p(wt), p(sh), p(out)) == fc._lib.ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fpga-cnn-object-detection-accelerator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_window_plan_for_large_images(lib):
    """csrc/tiling.cuh: every output owned by exactly one 128-pixel window, never by a window's outermost row/column unless
    that is the image border (where the kernel's zero padding is the real padding)."""
    buf = lambda: np.zeros(80, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for n_out in list(range(16, 130)) + [256, 1000, 1024]:
        g, s, e = buf(), buf(), buf()
        n = lib.cnnacc_tile_plan_host(n_out, p(g), p(s), p(e), 80)
        assert n >= 1
        owned = np.zeros(n_out, np.int32)
        for i in range(n):
            assert 0 <= g[i] <= n_out - 16 and s[i] < e[i]
            if n_out % 2 == 0:                       # H, W are multiples of 16, so n_out is even in practice: even window
                assert g[i] % 2 == 0                 # origins keep the TMA box's x coordinate 16-byte aligned (8*g - 16)
            for o in range(s[i], e[i]):
                loc = o - g[i]
                assert 0 <= loc <= 15 and (loc >= 1 or g[i] == 0) and (loc <= 14 or g[i] == n_out - 16)
                owned[o] += 1
        assert (owned == 1).all()
    assert lib.cnnacc_tile_plan_host(8, p(buf()), p(buf()), p(buf()), 80) == fc._lib.ERR_ARG
