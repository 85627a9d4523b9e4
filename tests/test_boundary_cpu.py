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
