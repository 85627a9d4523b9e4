"""CPU re-derivation of the fused kernel's operand layouts (no GPU needed).

The fused kernel (csrc/conv_fused.cuh) never materialises im2col: it points tcgen05 shared-memory descriptors
(start address, LBO, SBO) into halo-padded activation maps and into weights permuted by
cnnacc_pack_weights_host.  This test replays exactly that addressing in numpy -- a K-major no-swizzle operand
element (row, k) lives at  start + (row/8)*SBO + (k/16)*LBO + (row%8)*16 + k%16  -- builds the accumulators the
MMAs would produce, applies the kernel's epilogue (pool raw s32, then shift/saturate) and checks every layer
against the oracle.  It pins the packing (parse_kernels, arm_cnn.c:43-59, hoisted to load time) and the
descriptor arithmetic before any GPU time is spent.
"""
import ctypes
import os

import numpy as np
import pytest

import fpga_cnn_b200 as fc
import inputs
from oracle import np_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# constants mirrored from conv_fused.cuh
A1Q, A1P = 33 * 16, 2 * 33 * 16
A2Q, A2P, A2C = 17 * 16, 2 * 17 * 16, 34 * 2 * 17 * 16
B1_SLAB, B1_HALF = 4096, 2048


def pack(weights):
    lib = fc.load()
    w0 = np.zeros(352, np.uint32)
    b1 = np.zeros(24576, np.uint8)
    b2 = np.zeros(18432, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.cnnacc_pack_weights_host(p(weights), weights.size, p(w0), p(b1), p(b2)) == 0
    return w0[:96].reshape(16, 6), w0[96:].reshape(8, 32), b1, b2


def operand(buf, start, lbo, sbo, rows, signed):
    """Gather a (rows x 32) K-major no-swizzle operand through its descriptor."""
    r = np.arange(rows)[:, None]
    k = np.arange(32)[None, :]
    idx = start + (r // 8) * sbo + (k // 16) * lbo + (r % 8) * 16 + (k % 16)
    v = buf[idx]
    return (v.view(np.int8) if signed else v).astype(np.int64)


def act(v, shift):
    return np.clip(v >> shift, 0, 255).astype(np.uint8)


def layer0_dp4a(img, w0, shift):
    """Layer 0 as the dp4a loop computes it: 4-byte row windows dotted with lo / hi weight words."""
    pad = np.zeros((130, 160), np.int64)
    pad[1:129, 16:144] = img
    w = np.zeros((16, 6, 4), np.int64)
    for b in range(4):
        w[:, :, b] = ((w0 >> (8 * b)) & 0xFF).astype(np.uint8).view(np.int8).reshape(16, 6)
    out = np.zeros((16, 64, 64), np.uint8)
    yp, xp = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    cb = 2 * xp + 15
    A = np.stack([np.stack([pad[2 * yp + r, cb + b] for b in range(4)], -1) for r in range(4)])   # [4][64][64][4]
    for o in range(16):
        a00 = sum((A[d] * w[o, d]).sum(-1) for d in range(3))
        a01 = sum((A[d] * w[o, 3 + d]).sum(-1) for d in range(3))
        a10 = sum((A[d + 1] * w[o, d]).sum(-1) for d in range(3))
        a11 = sum((A[d + 1] * w[o, 3 + d]).sum(-1) for d in range(3))
        out[o] = act(np.maximum(np.maximum(a00, a01), np.maximum(a10, a11)), shift)
    return out


def layer0_imma(img, w0f, shift):
    """Layer 0 as the mma.sync fragments compute it: A row = a window's 4x4 patch (k = 4*row + col), B fragment word
    of lane 4*n + r in block (py, ol) = patch row r of output column n -> (oc = 4*(n>>1) + ol, px = n&1)."""
    pad = np.zeros((130, 160), np.int64)
    pad[1:129, 16:144] = img
    yp, xp = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    patch = np.stack([np.stack([pad[2 * yp + r, 2 * xp + 15 + c] for c in range(4)], -1) for r in range(4)], -2)   # [64][64][r][c]
    B = np.zeros((8, 8, 4, 4), np.int64)                                  # [blk][n][r][c]
    for c in range(4):
        B[:, :, :, c] = ((w0f >> (8 * c)) & 0xFF).astype(np.uint8).view(np.int8).reshape(8, 8, 4)
    D = np.einsum("yxrc,bnrc->yxbn", patch, B)                            # [64][64][blk][n]
    out = np.zeros((16, 64, 64), np.uint8)
    for n2 in range(4):
        for ol in range(4):
            members = [D[:, :, py * 4 + ol, 2 * n2 + px] for py in range(2) for px in range(2)]
            out[4 * n2 + ol] = act(np.maximum.reduce(members), shift)
    return out


def store_act1(l0):
    """[16][64][64] -> the act1 smem image [66][2][33][16]."""
    a1 = np.zeros(66 * A1P, np.uint8)
    for y in range(64):
        for x in range(64):
            off = (y + 1) * A1P + ((x + 1) & 1) * A1Q + ((x + 1) >> 1) * 16
            a1[off:off + 16] = l0[:, y, x]
    return a1


def layer1_umma(a1, b1, shift):
    """8 tiles x 8 Toeplitz K-slabs, N = 128 = 4 window members x 32 oc; pool over the members of a TMEM lane.
    The kernel's issue order: patch rows 1, 2 as N = 128 MMAs (the first one overwrites the accumulator), then patch row 0
    as N = 64 into columns 0-63 and patch row 3 as N = 64 into columns 64-127."""
    out = np.zeros((32, 32, 32), np.uint8)
    for t in range(8):
        ty, tx = t >> 2, t & 3
        a0 = (32 * ty) * A1P + (8 * tx) * 16
        D = None
        for q in range(4):
            r, sx = 1 + (q >> 1), q & 1
            A = operand(a1, a0 + r * A1P + sx * 16, A1Q, 2 * A1P, 128, signed=False)
            B = operand(b1, q * B1_SLAB, 2048, 128, 128, signed=True)
            D = A @ B.T if q == 0 else D + A @ B.T
        for q in range(4):
            r, sx = (3 if q >> 1 else 0), q & 1
            A = operand(a1, a0 + r * A1P + sx * 16, A1Q, 2 * A1P, 128, signed=False)
            B = operand(b1, 4 * B1_SLAB + q * B1_HALF, 1024, 128, 64, signed=True)
            D[:, 64 * (q >> 1):64 * (q >> 1) + 64] += A @ B.T
        pooled = D.reshape(128, 4, 32).max(axis=1)                   # lane = window, columns = member*32 + oc
        L = np.arange(128)
        out[:, 16 * ty + (L >> 3), 8 * tx + (L & 7)] = act(pooled, shift).T
    return out


def store_act2(l1):
    a2 = np.zeros(2 * A2C, np.uint8)
    for y in range(32):
        for x in range(32):
            for g in range(2):
                off = g * A2C + (y + 1) * A2P + ((x + 1) & 1) * A2Q + ((x + 1) >> 1) * 16
                a2[off:off + 16] = l1[16 * g:16 * g + 16, y, x]
    return a2


def layer2_umma(a2, b2, shift):
    out = np.zeros((64, 16, 16), np.uint8)
    for s in range(2):
        j0 = s * 8
        acc = []
        for p in range(4):
            a, b = p >> 1, p & 1
            D = np.zeros((128, 64), np.int64)
            for t in range(9):
                dy, dx = t // 3, t % 3
                start = (a + dy) * A2P + ((b + dx) & 1) * A2Q + (j0 + ((b + dx) >> 1)) * 16
                A = operand(a2, start, A2C, 2 * A2P, 128, signed=False)
                B = operand(b2, t * 2048, 1024, 128, 64, signed=True)
                D += A @ B.T
            acc.append(D)
        pooled = np.maximum(np.maximum(acc[0], acc[1]), np.maximum(acc[2], acc[3]))
        L = np.arange(128)
        out[:, L >> 3, j0 + (L & 7)] = act(pooled, shift).T
    return out


@pytest.fixture(scope="module", autouse=True)
def _built():
    fc.build()


@pytest.mark.parametrize("wkind,shifts", [("shipped", (2, 4, 6)), ("shipped", (7, 10, 11)), (("rng", 5), (9, 12, 13)),
                                          (("rng", 6), (0, 0, 0))])
def test_packed_operands_reproduce_the_oracle(wkind, shifts):
    shipped = np.fromfile(os.path.join(ROOT, "tests", "golden", "weights.bin"), dtype=np.uint8)
    wt = inputs.make_weights(wkind, shipped)
    img = inputs.make_images(("rng", 42), 1)[0]
    want, want_l0, want_l1 = np_oracle.infer(img, np_oracle.unpack_weights(wt), shifts, return_all=True)
    w0, w0f, b1, b2 = pack(wt)
    l0 = layer0_dp4a(img, w0, shifts[0])
    assert np.array_equal(l0, want_l0), "layer-0 dp4a words"
    assert np.array_equal(layer0_imma(img, w0f, shifts[0]), want_l0), "layer-0 mma.sync B fragments"
    l1 = layer1_umma(store_act1(l0), b1, shifts[1])
    assert np.array_equal(l1, want_l1), "layer-1 Toeplitz operand / descriptors"
    l2 = layer2_umma(store_act2(l1), b2, shifts[2])
    assert np.array_equal(l2.reshape(64, 256), want), "layer-2 operand / descriptors"


def test_toeplitz_density():
    """56.25 % of the layer-1 B operand is structurally non-zero (36 of 64 (member, patch pixel) pairs)."""
    wt = np.full(23184, 1, np.uint8)
    _, _, b1, _ = pack(wt)
    assert np.count_nonzero(b1) == 36 * 32 * 16
