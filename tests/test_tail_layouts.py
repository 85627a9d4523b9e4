"""CPU replay of the index arithmetic and the exactness arguments inside csrc/tail.cuh (no GPU needed).

tail_back sorts 256 CAM values with a bitonic network spread over 64 threads x 4 registers (strides 1, 2 in registers, 4..64 by
warp shuffle, 128 through shared memory) and reads sorted[178], sorted[179] from thread 44; tail_front squeezes a ballot into a
channel mask with three shift-or steps and forms the product w * float(byte) with one FMA.  Each of these is replayed here with
the same index expressions, so a wrong lane mask / direction bit / selector shows up before any GPU time is spent.
"""
import numpy as np
import pytest


def bitonic_as_the_kernel_does(vals):
    """vals[256] -> the network of tail_back: element e = 4T + b lives in thread T, register b."""
    v = np.array(vals, dtype=np.float32).reshape(64, 4).copy()            # v[T][b]
    T = np.arange(64)
    k = 2
    while k <= 256:
        up = np.ones(64, bool) if k == 256 else ((T & (k >> 2)) == 0)      # k == 2 decided per pair below
        j = k >> 1
        while j > 0:
            if j == 128:                                                   # exchange through shared memory: partner thread T ^ 32
                o = v[T ^ 32]
                lower = (T & 32) == 0
                v = np.where(lower[:, None], np.minimum(v, o), np.maximum(v, o))
            elif j >= 4:                                                   # warp shuffle: partner lane T ^ (j >> 2), same warp
                assert (j >> 2) < 32
                o = v[T ^ (j >> 2)]
                keep_min = ((T & (j >> 2)) == 0) == up
                v = np.where(keep_min[:, None], np.minimum(v, o), np.maximum(v, o))
            else:                                                          # inside the thread: registers a and a | j
                for a in range(4):
                    if a & j:
                        continue
                    asc = np.full(64, (a & 2) == 0) if k == 2 else up
                    lo, hi = np.minimum(v[:, a], v[:, a | j]), np.maximum(v[:, a], v[:, a | j])
                    v[:, a], v[:, a | j] = np.where(asc, lo, hi), np.where(asc, hi, lo)
            j >>= 1
        k <<= 1
    return v.reshape(-1)


@pytest.mark.parametrize("seed", range(6))
def test_bitonic_network_of_tail_back_sorts(seed):
    rng = np.random.default_rng(seed)
    vals = rng.random(256).astype(np.float32)
    if seed % 2:                                                           # CAMs are full of ties: zeros after ReLU, ones at the max
        vals[rng.random(256) < 0.6] = 0.0
        vals[rng.integers(0, 256, 5)] = 1.0
    got = bitonic_as_the_kernel_does(vals)
    assert np.array_equal(got, np.sort(vals))
    # np.percentile(cam, 70) of 256 values interpolates between sorted[178] and sorted[179]: thread 44, registers 2 and 3
    assert 4 * 44 + 2 == 178 and 4 * 44 + 3 == 179
    lo, hi = got[178], got[179]
    thr = np.float32(hi - np.float32(np.float32(hi - lo) * np.float32(0.5)))
    assert abs(float(thr) - float(np.percentile(vals.astype(np.float64), 70))) <= 1e-6


def test_ballot_squeeze_and_channel_mapping():
    """Lanes 4q..4q+3 of warp w hold channel 32 i + 8 w + q (task T + 128 i = channel T/4 + 32 i): the ballot has four equal bits
    per channel, and three shift-or steps squeeze bits 0, 4, .., 28 into one byte that lands at bit 32 i + 8 w."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        valid = rng.integers(0, 2, 64).astype(bool)                        # per channel
        mask = 0
        for i in range(2):
            for w in range(4):
                ballot = 0
                for lane in range(32):
                    T = 32 * w + lane
                    ch = (T + 128 * i) // 4
                    assert ch == 32 * i + 8 * w + lane // 4
                    ballot |= int(valid[ch]) << lane
                b = ballot & 0x11111111
                b = (b | (b >> 3)) & 0x03030303
                b = (b | (b >> 6)) & 0x000F000F
                b = (b | (b >> 12)) & 0xFF
                mask |= b << (32 * i + 8 * w)
        assert mask == sum(int(v) << c for c, v in enumerate(valid))


def test_product_with_one_fma_equals_the_separately_rounded_product():
    """cam += w * float(byte) with product and sum rounded separately (numpy's reduction).  The kernel builds x = 2^23 + byte with
    one PRMT (0x4B0000bb) and takes fma(w, x, -w * 2^23): w * 2^23 is exact, w * x - w * 2^23 = w * byte exactly, and the FMA rounds
    that exact value once -- i.e. it IS fmul_rn(w, float(byte)).  Replayed in float64, where every intermediate is exact."""
    rng = np.random.default_rng(2)
    w = np.concatenate([rng.standard_normal(4000) * 0.1, rng.standard_normal(500) * 1e-30, rng.standard_normal(500) * 1e30,
                        [0.0, -0.0, 1.0, -1.0, np.float32(1e-40), np.float32(2.0 ** 99)]]).astype(np.float32)   # |w| < 2^100: load_classifier's bound
    for b in range(256):
        x = np.array([0x4B000000 | b], dtype=np.uint32).view(np.float32)[0]
        assert float(x) == 2.0 ** 23 + b
        c = w.astype(np.float64) * -(2.0 ** 23)                            # exact (a power of two), and finite in fp32 for these w
        assert np.all(np.isfinite(c.astype(np.float32)))
        exact = w.astype(np.float64) * float(x) + c                        # 48-bit product, exact sum = w * b (at most 32 bits)
        assert np.array_equal(exact, w.astype(np.float64) * b)
        with np.errstate(under="ignore"):
            fma = exact.astype(np.float32)                                 # the single rounding of the FMA
            mul = w * np.float32(b)                                        # fmul_rn(w, float(b))
        assert np.array_equal(fma.view(np.uint32) & 0x7FFFFFFF, mul.view(np.uint32) & 0x7FFFFFFF)   # up to the sign of zero
        nz = mul != 0
        assert np.array_equal(fma[nz], mul[nz])


def test_cam_thread_to_pixel_and_bin_mapping():
    """tail_front: thread T of 128 owns pixels 2T, 2T+1 (16-bit load at byte 2T of each channel plane) and reads the class weight of
    bin (T/32)*4 + (T%8)/2; tail_back: thread T of 64 owns pixels 4T..4T+3 = row T/4, columns 4(T%4)..+3."""
    for T in range(128):
        for b in range(2):
            p = 2 * T + b
            y, x = p // 16, p % 16
            assert y == T >> 3 and x == 2 * (T & 7) + b
            assert (y // 4) * 4 + x // 4 == (T >> 5) * 4 + ((T & 7) >> 1)
    for T in range(64):
        for b in range(4):
            p = 4 * T + b
            assert p // 16 == T >> 2 and p % 16 == 4 * (T & 3) + b
