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


def _pdl(launches, sm_count=148):
    """launches: list of (in_lo, in_hi, out_lo, out_hi, n_images, stream, foreign) -> list of wait decisions."""
    import ctypes
    lib = fc.load()
    n = len(launches)
    ranges = np.array([v for l in launches for v in l[:4]], dtype=np.uint64)
    n_img = np.array([l[4] for l in launches], dtype=np.int64)
    stream = np.array([l[5] for l in launches], dtype=np.int32)
    foreign = np.array([l[6] for l in launches], dtype=np.int32)
    out = np.full(n, -1, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.cnnacc_pdl_chain_host(n, p(ranges), p(n_img), p(stream), p(foreign), sm_count, p(out)) == 0
    return out.tolist()


def test_overlapped_launch_bookkeeping():
    """The rules behind `griddepcontrol.wait` being skipped (csrc/pdl_chain.h): a launch may overlap its predecessor only if
    it follows it directly on the same stream, fills the GPU, and shares no memory (RAW, WAW, WAR) with any launch since the
    last one that waited."""
    K = 16384
    buf = lambda i, n=4096: (0x1000_0000 + i * 0x1000_0000, 0x1000_0000 + i * 0x1000_0000 + n * K)
    L = lambda i_in, i_out, n=4096, stream=7, foreign=0: (*buf(i_in, n), *buf(i_out, n), n, stream, foreign)
    # 1. rotating independent buffers: the first launch waits, the rest overlap, a full chain (16) forces a wait
    seq = [L(2 * i, 2 * i + 1) for i in range(20)]
    w = _pdl(seq)
    assert w[0] == 1 and w[1:16] == [0] * 15 and w[16] == 1 and w[17:] == [0] * 3
    # 2. WAW: same output buffer twice -> wait
    assert _pdl([L(0, 1), L(2, 1)]) == [1, 1]
    # 3. RAW: reads what the previous launch wrote -> wait; also against an OLDER member of the chain
    assert _pdl([L(0, 1), L(1, 2)]) == [1, 1]
    assert _pdl([L(0, 1), L(2, 3), L(1, 4)]) == [1, 0, 1]
    # 4. WAR: writes what a chain member still reads -> wait
    assert _pdl([L(0, 1), L(2, 0)]) == [1, 1]
    # 5. partial overlap of ranges counts
    a = (0x1000, 0x1000 + 4096 * K, 0x9000_0000, 0x9000_0000 + 4096 * K, 4096, 7, 0)
    b = (0x5000_0000, 0x5000_0000 + 4096 * K, 0x9000_0000 + 4096 * K - 1, 0x9000_0000 + 2 * 4096 * K, 4096, 7, 0)
    assert _pdl([a, b]) == [1, 1]
    # 6. a grid smaller than the SM count always waits and is no fence: its successor waits too
    assert _pdl([L(0, 1), L(2, 3), L(4, 5, n=100), L(6, 7), L(8, 9)]) == [1, 0, 1, 1, 0]
    # 7. another kernel of the handle, another handle's conv launch, or another stream in between -> wait
    assert _pdl([L(0, 1), L(2, 3, foreign=1), L(4, 5)]) == [1, 1, 0]
    assert _pdl([L(0, 1), L(2, 3, foreign=2), L(4, 5)]) == [1, 1, 0]
    assert _pdl([L(0, 1), L(2, 3, stream=8), L(4, 5, stream=8)]) == [1, 1, 0]
    # 8. after a waiting launch only the launches since then matter: buffer 1 is free to be reused
    assert _pdl([L(0, 1), L(2, 3), L(4, 3), L(6, 1)]) == [1, 0, 1, 0]


def _chunk_plan(lib, n, H=128, W=128, pipelined=0):
    sizes = np.zeros(4096, np.int64)
    k = lib.cnnacc_chunk_plan_host(n, H, W, pipelined, sizes.ctypes.data_as(ctypes.c_void_p), sizes.size)
    assert k > 0, (n, H, W, pipelined, k)
    return [int(v) for v in sizes[:k]]


def test_host_chunk_plans_cover_every_call_once(lib):
    """How host-pointer calls are cut into staging chunks (csrc/host_chunks.h): every plan covers the call exactly once with
    chunks of 1..full images; synchronous calls ramp 1/4, 1/2, 1 ... 1, 1/2, 1/4; streamed calls use one chunk up to 64 MiB."""
    # the bench's e2e batch, by hand: 64 MiB call -> 16 MiB chunks (1024 images), ramp depth 2
    assert _chunk_plan(lib, 4096) == [256, 512, 1024, 1024, 512, 512, 256]
    assert _chunk_plan(lib, 4096, pipelined=1) == [4096]
    assert _chunk_plan(lib, 16384, pipelined=1) == [4096] * 4
    assert _chunk_plan(lib, 16384 + 5, pipelined=1) == [4096] * 4 + [5]
    assert _chunk_plan(lib, 1) == [1] and _chunk_plan(lib, 1, pipelined=1) == [1]
    assert _chunk_plan(lib, 65) == [65]                                   # < 4 MiB: one chunk
    rng = np.random.default_rng(0)
    cases = [(int(n), 128, 128) for n in list(range(1, 40)) + [255, 256, 257, 1023, 1024, 1025, 4095, 4097, 65536, 1 << 20]]
    cases += [(int(rng.integers(1, 300000)), 128, 128) for _ in range(300)]
    cases += [(int(rng.integers(1, 3000)), 16 * int(rng.integers(1, 40)), 16 * int(rng.integers(1, 40))) for _ in range(200)]
    for n, H, W in cases:
        in_sz = H * W
        for pipelined in (0, 1):
            plan = _chunk_plan(lib, n, H, W, pipelined)
            assert sum(plan) == n and min(plan) >= 1, (n, H, W, pipelined)
            full = max(plan)
            cap = (64 << 20) if pipelined else (32 << 20)
            assert full * in_sz <= max(cap, in_sz) and full * in_sz * 8 <= max(512 << 20, in_sz * 8), (n, H, W, pipelined)
            if pipelined:
                assert all(m == full for m in plan[:-1]) and len(plan) == -(-n // full), (n, H, W)
            else:
                # ramp: sizes never decrease up to the first full-size chunk; after the last one comes at most one odd-sized
                # remainder piece and then the halving ramp-down
                first, last = plan.index(full), len(plan) - 1 - plan[::-1].index(full)
                down = plan[last + 1:]
                if len(down) > 1 and down[0] < down[1]:
                    down = down[1:]
                assert plan[:first + 1] == sorted(plan[:first + 1]) and down == sorted(down, reverse=True), (n, H, W, plan)
                assert len(plan) <= n // full + 6, (n, H, W, plan)
    assert lib.cnnacc_chunk_plan_host(0, 128, 128, 0, None, 0) == fc._lib.ERR_ARG
