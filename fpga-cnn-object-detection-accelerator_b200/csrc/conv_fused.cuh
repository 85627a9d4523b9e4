// conv_fused.cuh -- the hot path: one persistent sm_100a kernel for the whole 128x128 conv stack.
//
// What it replaces: cnn_infer (/root/reference/software/arm_cnn.c:159-198) == the PL datapath
// layer_fsm + conv_core + accumulator + ReLU + max_pooling_engine over feature/weight BRAM
// (rtl/core/cnn_acc_top.v).  Like the FPGA design, every intermediate map stays on chip: one CTA per SM
// keeps an image's maps in shared memory and HBM sees 16 KiB of pixels in and 16 KiB of features out.
//
//   layer 0  (1->16, 128x128, K=9)    CUDA cores: dp4a.u32.s32, 2x2 pool in registers, -> act1 (smem)
//   layer 1  (16->32, 64x64, K=144)   tcgen05.mma kind::i8 (A = u8 activations, B = s8 weights, D = s32 in TMEM)
//   layer 2  (32->64, 32x32, K=288)   tcgen05.mma kind::i8
//   each followed by >>shift, ReLU/saturate (arm_cnn.c:127-135) and 2x2 max-pool (arm_cnn.c:115-143),
//   pooled on the raw s32 first (monotone activation).
//
// Implicit GEMM without im2col.  Activation maps are stored as [y+1][x-parity][(x+1)/2][16 ch] bytes with a
// zero halo, so a no-swizzle K-major UMMA core matrix (8 rows x 16 B) is "8 same-parity pixels x 16 channels",
// a conv tap is a 16-byte-granular start-address offset, and SBO = 2 row pitches makes the 128 rows of an MMA
// a 16-row-pair x 8-col-pair block of ONE output parity (y%2, x%2).  The four parities of a block go to four
// TMEM column groups, so TMEM lane l holds all four members of pooling window l: the pool is thread-local.
//   layer 1: K=32 per MMA = two taps (LBO = distance between them): (0,dx)+(1,dx) for dx=0..2, (2,0)+(2,2),
//            (2,1)+zeros -> 5 MMAs per parity tile, N = 32.
//   layer 2: K=32 per MMA = one tap over both 16-channel planes (LBO = plane stride) -> 9 MMAs, N = 64.
// These descriptor forms were verified on a B200 by tools/probe_umma.cu (profiles/r1_probe_umma_dp4a_tmem.txt).
//
// Warps: 16 compute warps (layer-0 conv, then TMEM epilogues: warp%4 = TMEM lane quarter, warp/4 = channel
// group) + 1 control warp (TMA loads of the next images, MMA issue by one elected thread).
#pragma once
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "weights_pack.h"

namespace cnnacc {

// ---- shared-memory plan (bytes) ---------------------------------------------------------------------
constexpr int kInPitch   = 160;                       // x = -16 .. 143 (TMA box, OOB zero-filled; the innermost
                                                      // box coordinate must be 16-byte aligned: tools/probe_tma.cu)
constexpr int kInRows    = 130;                       // y = -1 .. 128
constexpr int kInBytes   = kInPitch * kInRows;        // 20800, one TMA transaction
constexpr int kInStride  = 20864;                     // 128-byte aligned slot size
constexpr int kA1Q       = 33 * 16;                   // act1 parity-plane stride   (528)
constexpr int kA1P       = 2 * kA1Q;                  // act1 row pitch             (1056)
constexpr int kA1Bytes   = 66 * kA1P;                 // 69696
constexpr int kA2Q       = 17 * 16;                   // act2 parity-plane stride   (272)
constexpr int kA2P       = 2 * kA2Q;                  // act2 row pitch             (544)
constexpr int kA2C       = 34 * kA2P;                 // act2 channel-block plane   (18496)
constexpr int kA2Bytes   = 2 * kA2C;                  // 36992
constexpr int kB1Bytes   = 5 * 1024;                  // layer-1 B: 5 MMAs x (2 K-halves x 4 row groups x 128 B)
constexpr int kB2Bytes   = 9 * 2048;                  // layer-2 B: 9 taps x (2 K-halves x 8 row groups x 128 B)

constexpr int kOffIn0  = 0;
constexpr int kOffIn1  = kInStride;
constexpr int kOffA1   = 2 * kInStride;               // 41728
constexpr int kOffA2   = kOffA1 + 69760;              // 107392
constexpr int kOffB1   = kOffA2 + kA2Bytes;           // 144384
constexpr int kOffB2   = kOffB1 + kB1Bytes;           // 149504
constexpr int kOffBar  = kOffB2 + kB2Bytes;           // 167936
constexpr int kFusedSmem = kOffBar + 128;

constexpr int kComputeWarps = 16;
constexpr int kFusedThreads = (kComputeWarps + 1) * 32;    // 544
constexpr uint32_t kTmemCols = 512;

// error bits reported through the status word
constexpr int kErrInputTimeout = 1, kErrMmaTimeout = 2, kErrEmptyTimeout = 4, kErrWeightTimeout = 8;

struct FusedParams {
    uint32_t w0[16][6];          // layer-0 dp4a words per out-channel: lo[dy], hi[dy]  (constant bank)
    int shift0, shift1, shift2;
    int n_images;
    const uint8_t* b1;           // packed layer-1 B operand (kB1Bytes)
    const uint8_t* b2;           // packed layer-2 B operand (kB2Bytes)
    uint8_t* out;                // [n][64][16][16]
    uint8_t* dump_l0;            // optional [n][16][64][64]
    uint8_t* dump_l1;            // optional [n][32][32][32]
    int* status;                 // device int, OR-ed error bits
    int debug_level;             // bring-up bisection: run only the first stages (99 = everything)
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a broken pipeline must never hang the GPU.  Returns false on timeout.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, long long budget) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (clock64() - t0 > budget) return false;
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (version 1 = sm_100).  Offsets in bytes, multiples of 16.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 |
           (uint64_t)((sbo >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46;
}
// kind::i8 instruction descriptor: D = s32, A = unsigned 8-bit, B = signed 8-bit, both K-major, M = 128.
__device__ __forceinline__ constexpr uint32_t umma_idesc_i8(int n) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 3D tiled TMA load (x, y, image) -> smem, completion on an mbarrier.
__device__ __forceinline__ void tma_load_image(uint32_t dst, const CUtensorMap* map, uint32_t bar, int img) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(-16), "r"(-1), "r"(img), "r"(bar) : "memory");
}
// 1D bulk copy global -> smem (pre-packed weights).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- the kernel ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFusedThreads, 1)
conv_stack_fused_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ FusedParams P)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s_base = smem_u32(smem);
    // barriers: 0,1 input slot full; 2,3 TMEM buffer full (MMA committed); 4,5 TMEM buffer empty; 6 weights
    const uint32_t bar_in = s_base + kOffBar, bar_full = bar_in + 16, bar_empty = bar_in + 32, bar_w = bar_in + 48;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 64);
    int* s_err = reinterpret_cast<int*>(smem + kOffBar + 72);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_ctrl = (warp == kComputeWarps);
    const int n_local = (P.n_images - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // images of this CTA

    // ---- one-time setup ---------------------------------------------------------------------------------
    for (int i = tid; i < (kA1Bytes + 64 + kA2Bytes) / 16; i += kFusedThreads)                 // zero halos (and interiors)
        reinterpret_cast<uint4*>(smem + kOffA1)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar_in, 1); mbar_init(bar_in + 8, 1);
        mbar_init(bar_full, 1); mbar_init(bar_full + 8, 1);
        mbar_init(bar_empty, kComputeWarps); mbar_init(bar_empty + 8, kComputeWarps);
        mbar_init(bar_w, 1);
        *s_err = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (is_ctrl) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;

    const int dbg = P.debug_level;
    if (is_ctrl && lane == 0 && dbg >= 2) {
        mbar_expect_tx(bar_w, kB1Bytes + kB2Bytes);
        bulk_load(s_base + kOffB1, P.b1, kB1Bytes, bar_w);
        bulk_load(s_base + kOffB2, P.b2, kB2Bytes, bar_w);
        for (int k = 0; k < 2 && k < n_local && dbg >= 3; k++) {
            mbar_expect_tx(bar_in + 8 * k, kInBytes);
            tma_load_image(s_base + (k ? kOffIn1 : kOffIn0), &in_map, bar_in + 8 * k, (int)blockIdx.x + k * (int)gridDim.x);
        }
    }

    long long budget = 200000000LL;                      // ~0.1 s; collapses after the first timeout
    auto wait_or_flag = [&](uint32_t bar, uint32_t parity, int code) {
        if (!mbar_wait(bar, parity, *reinterpret_cast<volatile int*>(s_err) ? 2000LL : budget)) {
            atomicOr(s_err, code);
        }
    };
    if (is_ctrl && lane == 0 && dbg >= 2) wait_or_flag(bar_w, 0, kErrWeightTimeout);

    uint32_t full_uses[2] = {0, 0};                      // per TMEM buffer: completed uses (same sequence in every role)

    for (int k = 0; k < n_local && dbg >= 3; k++) {
        const int img = (int)blockIdx.x + k * (int)gridDim.x;
        const int slot = k & 1;

        // =============== layer 0: dp4a on CUDA cores ========================================================
        if (!is_ctrl) {
            wait_or_flag(bar_in + 8 * slot, (uint32_t)(k >> 1) & 1, kErrInputTimeout);
            const uint32_t* in_w = reinterpret_cast<const uint32_t*>(smem + (slot ? kOffIn1 : kOffIn0));
#pragma unroll 1
            for (int it = 0; it < (dbg >= 4 ? 8 : 0); it++) {
                const int pidx = it * 512 + tid;
                const int yp = pidx >> 6;
                const int xp = (pidx & 32) + 2 * (lane & 15) + (lane >> 4);      // lanes 0-15 even x, 16-31 odd x
                const int cb = 2 * xp + 15;                                      // smem byte of pixel column 2xp-1
                const uint32_t* rp = in_w + (2 * yp) * (kInPitch / 4) + (cb >> 2);
                const int sh = (cb & 3) * 8;
                uint32_t A[4];
#pragma unroll
                for (int r = 0; r < 4; r++) A[r] = __funnelshift_r(rp[r * (kInPitch / 4)], rp[r * (kInPitch / 4) + 1], sh);
                int pooled[16];
#pragma unroll
                for (int o = 0; o < 16; o++) {
                    const uint32_t l0 = P.w0[o][0], l1 = P.w0[o][1], l2 = P.w0[o][2];
                    const uint32_t h0 = P.w0[o][3], h1 = P.w0[o][4], h2 = P.w0[o][5];
                    int a00 = dp4a_u8s8(A[0], l0, dp4a_u8s8(A[1], l1, dp4a_u8s8(A[2], l2, 0)));
                    int a01 = dp4a_u8s8(A[0], h0, dp4a_u8s8(A[1], h1, dp4a_u8s8(A[2], h2, 0)));
                    int a10 = dp4a_u8s8(A[1], l0, dp4a_u8s8(A[2], l1, dp4a_u8s8(A[3], l2, 0)));
                    int a11 = dp4a_u8s8(A[1], h0, dp4a_u8s8(A[2], h1, dp4a_u8s8(A[3], h2, 0)));
                    pooled[o] = max4(a00, a01, a10, a11);
                }
                uint4 v;
                v.x = act_pack4(pooled[0], pooled[1], pooled[2], pooled[3], P.shift0);
                v.y = act_pack4(pooled[4], pooled[5], pooled[6], pooled[7], P.shift0);
                v.z = act_pack4(pooled[8], pooled[9], pooled[10], pooled[11], P.shift0);
                v.w = act_pack4(pooled[12], pooled[13], pooled[14], pooled[15], P.shift0);
                *reinterpret_cast<uint4*>(smem + kOffA1 + (yp + 1) * kA1P + ((xp + 1) & 1) * kA1Q + ((xp + 1) >> 1) * 16) = v;
            }
            fence_async_smem();                          // act1 (generic proxy) -> visible to the MMA (async proxy)
        }
        __syncthreads();                                 // act1 complete; input slot free

        if (P.dump_l0) {                                 // debug / register-protocol path: BRAM channels 0-15
            for (int i = tid; i < 16 * 4096; i += kFusedThreads) {
                const int c = i >> 12, y = (i >> 6) & 63, x = i & 63;
                P.dump_l0[(size_t)img * 65536 + i] = smem[kOffA1 + (y + 1) * kA1P + ((x + 1) & 1) * kA1Q + ((x + 1) >> 1) * 16 + c];
            }
        }

        if (dbg < 5) continue;
        if (is_ctrl) {
            if (lane == 0) {
                // prefetch the image after next into the slot layer 0 just released
                if (k + 2 < n_local) {
                    fence_async_smem();
                    mbar_expect_tx(bar_in + 8 * slot, kInBytes);
                    tma_load_image(s_base + (slot ? kOffIn1 : kOffIn0), &in_map, bar_in + 8 * slot, img + 2 * (int)gridDim.x);
                }
                // =============== layer 1 MMAs: 8 blocks x 4 parities x 5 K-steps, N = 32 ============================
                tc_fence_after();
                constexpr uint32_t idesc1 = umma_idesc_i8(32);
#pragma unroll 1
                for (int s = 0; s < 8; s++) {
                    const int buf = s & 1, i0 = (s >> 2) * 16, j0 = (s & 3) * 8;
                    wait_or_flag(bar_empty + 8 * buf, (full_uses[buf] & 1) ^ 1, kErrEmptyTimeout);
                    tc_fence_after();
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const int a = p >> 1, b = p & 1;
                        const uint32_t d = tm + buf * 128 + p * 32;
#pragma unroll
                        for (int m = 0; m < 5; m++) {
                            const int dy = (m < 3) ? 0 : 2, dx = (m < 3) ? m : (m == 3 ? 0 : 1);
                            const uint32_t lbo = (m < 3) ? kA1P : (m == 3 ? 16 : 0);
                            const uint32_t aaddr = s_base + kOffA1 + (2 * i0 + a + dy) * kA1P + ((b + dx) & 1) * kA1Q + (j0 + ((b + dx) >> 1)) * 16;
                            umma_i8(d, umma_desc(aaddr, lbo, 2 * kA1P), umma_desc(s_base + kOffB1 + m * 1024, 512, 128), idesc1, m > 0);
                        }
                    }
                    umma_commit(bar_full + 8 * buf);
                    full_uses[buf]++;
                }
            }
            __syncwarp();
        } else {
            // =============== layer 1 epilogue: TMEM -> pool -> shift/ReLU/saturate -> act2 (smem) ================
            const int q = warp & 3, g = warp >> 2;       // TMEM lane quarter, group of 8 output channels
            const int L = q * 32 + lane;
#pragma unroll 1
            for (int s = 0; s < 8; s++) {
                const int buf = s & 1, i0 = (s >> 2) * 16, j0 = (s & 3) * 8;
                wait_or_flag(bar_full + 8 * buf, full_uses[buf] & 1, kErrMmaTimeout);
                full_uses[buf]++;
                tc_fence_after();
                int v[4][8];
                const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + buf * 128 + g * 8;
#pragma unroll
                for (int p = 0; p < 4; p++) tmem_ld8(taddr + p * 32, v[p]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
                int m[8];
#pragma unroll
                for (int c = 0; c < 8; c++) m[c] = max4(v[0][c], v[1][c], v[2][c], v[3][c]);
                uint2 w;
                w.x = act_pack4(m[0], m[1], m[2], m[3], P.shift1);
                w.y = act_pack4(m[4], m[5], m[6], m[7], P.shift1);
                const int i = i0 + (L >> 3), j = j0 + (L & 7);
                *reinterpret_cast<uint2*>(smem + kOffA2 + (g >> 1) * kA2C + (i + 1) * kA2P + ((j + 1) & 1) * kA2Q +
                                          ((j + 1) >> 1) * 16 + (g & 1) * 8) = w;
            }
            fence_async_smem();
        }
        tc_fence_before();
        __syncthreads();                                 // act2 complete
        tc_fence_after();

        if (P.dump_l1) {                                 // BRAM channels 16-47
            for (int i = tid; i < 32 * 1024; i += kFusedThreads) {
                const int c = i >> 10, y = (i >> 5) & 31, x = i & 31;
                P.dump_l1[(size_t)img * 32768 + i] =
                    smem[kOffA2 + (c >> 4) * kA2C + (y + 1) * kA2P + ((x + 1) & 1) * kA2Q + ((x + 1) >> 1) * 16 + (c & 15)];
            }
        }

        if (dbg < 6) continue;
        if (is_ctrl) {
            if (lane == 0) {
                // =============== layer 2 MMAs: 2 blocks x 4 parities x 9 taps, N = 64 ==============================
                constexpr uint32_t idesc2 = umma_idesc_i8(64);
#pragma unroll 1
                for (int s = 0; s < 2; s++) {
                    const int buf = s, j0 = s * 8;
                    wait_or_flag(bar_empty + 8 * buf, (full_uses[buf] & 1) ^ 1, kErrEmptyTimeout);
                    tc_fence_after();
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        const int a = p >> 1, b = p & 1;
                        const uint32_t d = tm + buf * 256 + p * 64;
#pragma unroll
                        for (int t = 0; t < 9; t++) {
                            const int dy = t / 3, dx = t % 3;
                            const uint32_t aaddr = s_base + kOffA2 + (a + dy) * kA2P + ((b + dx) & 1) * kA2Q + (j0 + ((b + dx) >> 1)) * 16;
                            umma_i8(d, umma_desc(aaddr, kA2C, 2 * kA2P), umma_desc(s_base + kOffB2 + t * 2048, 1024, 128), idesc2, t > 0);
                        }
                    }
                    umma_commit(bar_full + 8 * buf);
                    full_uses[buf]++;
                }
            }
            __syncwarp();
        } else {
            // =============== layer 2 epilogue: TMEM -> pool -> activation -> features (HBM, CHW) ==================
            const int q = warp & 3, g = warp >> 2;       // group of 16 output channels
            const int L = q * 32 + lane;
            uint8_t* out_img = P.out + (size_t)img * 16384;
#pragma unroll 1
            for (int s = 0; s < 2; s++) {
                const int buf = s, j0 = s * 8;
                wait_or_flag(bar_full + 8 * buf, full_uses[buf] & 1, kErrMmaTimeout);
                full_uses[buf]++;
                tc_fence_after();
                const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + buf * 256 + g * 16;
                int m[16];
                {
                    int v0[16], v1[16];
                    tmem_ld16(taddr, v0);
                    tmem_ld16(taddr + 64, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; c++) m[c] = max(v0[c], v1[c]);
                    tmem_ld16(taddr + 128, v0);
                    tmem_ld16(taddr + 192, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; c++) m[c] = max(m[c], max(v0[c], v1[c]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
                const int i = L >> 3, j = j0 + (L & 7);
                uint8_t* o = out_img + (g * 16) * 256 + i * 16 + j;
#pragma unroll
                for (int c = 0; c < 16; c++) o[c * 256] = (uint8_t)min(max(m[c], 0) >> P.shift2, 255);
            }
        }
        // the next image's layer 0 only touches the input slot and act1; act1's last readers (layer-1 MMAs)
        // completed before the layer-1 epilogue finished, so no further barrier is needed here.
    }

    // ---- teardown ---------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && *s_err) atomicOr(P.status, *s_err);
    if (is_ctrl) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(kTmemCols) : "memory");
}

// ---- host side ------------------------------------------------------------------------------------------------
struct FusedWeights {
    bool ready = false;
    uint32_t w0[16][6];
    uint8_t* d_b1 = nullptr;
    uint8_t* d_b2 = nullptr;
    int* d_status = nullptr;
    bool attr_set = false;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Permute weights.bin into the operand layouts above.  Returns a cudaError_t as int.
inline int fused_load_weights(FusedWeights& fw, const uint8_t* wbin) {
    fw.ready = false;
    for (int o = 0; o < 16; o++)
        for (int dy = 0; dy < 3; dy++) {
            uint32_t lo = (uint32_t)weight_byte(wbin, 0, o, 0, dy * 3) | (uint32_t)weight_byte(wbin, 0, o, 0, dy * 3 + 1) << 8 |
                          (uint32_t)weight_byte(wbin, 0, o, 0, dy * 3 + 2) << 16;
            fw.w0[o][dy] = lo;
            fw.w0[o][3 + dy] = lo << 8;
        }
    // layer 1: MMA m pairs taps (first, second) in K bytes 0-15 / 16-31; B[n][k] at m*1024 + (k/16)*512 + (n/8)*128 + (n%8)*16 + k%16
    static const int pair_tap[5][2] = {{0, 3}, {1, 4}, {2, 5}, {6, 8}, {7, -1}};
    std::vector<uint8_t> b1(kB1Bytes, 0), b2(kB2Bytes, 0);
    for (int m = 0; m < 5; m++)
        for (int kc = 0; kc < 2; kc++) {
            const int tap = pair_tap[m][kc];
            if (tap < 0) continue;
            for (int n = 0; n < 32; n++)
                for (int ic = 0; ic < 16; ic++)
                    b1[m * 1024 + kc * 512 + (n / 8) * 128 + (n % 8) * 16 + ic] = weight_byte(wbin, 1, n, ic, tap);
        }
    // layer 2: tap t, K = input channel; B[n][k] at t*2048 + (k/16)*1024 + (n/8)*128 + (n%8)*16 + k%16
    for (int t = 0; t < 9; t++)
        for (int n = 0; n < 64; n++)
            for (int ic = 0; ic < 32; ic++)
                b2[t * 2048 + (ic / 16) * 1024 + (n / 8) * 128 + (n % 8) * 16 + (ic % 16)] = weight_byte(wbin, 2, n, ic, t);
    cudaError_t e;
    if (!fw.d_b1 && (e = cudaMalloc(&fw.d_b1, kB1Bytes)) != cudaSuccess) return (int)e;
    if (!fw.d_b2 && (e = cudaMalloc(&fw.d_b2, kB2Bytes)) != cudaSuccess) return (int)e;
    if (!fw.d_status) {
        if ((e = cudaMalloc(&fw.d_status, sizeof(int))) != cudaSuccess) return (int)e;
        if ((e = cudaMemset(fw.d_status, 0, sizeof(int))) != cudaSuccess) return (int)e;
    }
    if ((e = cudaMemcpy(fw.d_b1, b1.data(), kB1Bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpy(fw.d_b2, b2.data(), kB2Bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    if (!fw.attr_set) {
        if ((e = cudaFuncSetAttribute(conv_stack_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem)) != cudaSuccess) return (int)e;
        fw.attr_set = true;
    }
    if (!get_encode_tiled()) return (int)cudaErrorNotSupported;
    fw.ready = true;
    return 0;
}

inline void fused_free(FusedWeights& fw) {
    cudaFree(fw.d_b1); cudaFree(fw.d_b2); cudaFree(fw.d_status);
    fw.d_b1 = fw.d_b2 = nullptr; fw.d_status = nullptr; fw.ready = false;
}

// One launch for n device-resident images.  Returns a cudaError_t as int (0 = launched).
inline int launch_fused(const FusedWeights& fw, cudaStream_t stream, const uint8_t* d_imgs, int64_t n, uint8_t* d_feats,
                        const int* shifts, int sm_count, uint8_t* dump_l0, uint8_t* dump_l1) {
    if (n <= 0) return 0;
    if (n > 0x7fffffff || (reinterpret_cast<uintptr_t>(d_imgs) & 15)) return (int)cudaErrorInvalidValue;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {128, 128, (cuuint64_t)n};
    const cuuint64_t gstride[2] = {128, 16384};
    const cuuint32_t box[3] = {(cuuint32_t)kInPitch, (cuuint32_t)kInRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_imgs), gdim, gstride, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    FusedParams P;
    std::memcpy(P.w0, fw.w0, sizeof(P.w0));
    P.shift0 = shifts[0]; P.shift1 = shifts[1]; P.shift2 = shifts[2];
    P.n_images = (int)n;
    P.b1 = fw.d_b1; P.b2 = fw.d_b2;
    P.out = d_feats; P.dump_l0 = dump_l0; P.dump_l1 = dump_l1;
    P.status = fw.d_status;
    { const char* e = getenv("CNNACC_DEBUG_LEVEL"); P.debug_level = e ? atoi(e) : 99; }
    const int grid = (int)std::min<int64_t>(n, sm_count);
    conv_stack_fused_kernel<<<grid, kFusedThreads, kFusedSmem, stream>>>(map, P);
    return (int)cudaGetLastError();
}

// Reads (and clears) the device status word; non-zero = a pipeline wait timed out inside some launch.
inline int fused_poll_status(const FusedWeights& fw, int* bits) {
    *bits = 0;
    if (!fw.d_status) return 0;
    cudaError_t e = cudaMemcpy(bits, fw.d_status, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    if (*bits) e = cudaMemset(fw.d_status, 0, sizeof(int));
    return (int)e;
}

}  // namespace cnnacc
