// int8_peak_probe.cuh -- measures the dense int8 tensor-core ceiling of THIS GPU in the run that quotes it.
//
// MEASURED_PEAKS.json carries a bf16 GEMM and a copy bandwidth but no int8 number, and the conv stack's roofline is the
// tcgen05 kind::i8 rate.  One warp per SM issues back-to-back tcgen05.mma.cta_group::1.kind::i8 with M = 128, N = 256,
// K = 32 (the shape with the highest rate, profiles/r1_probe_umma_rate.txt: 128 clk per MMA = 8192 MAC/clk/SM) from
// shared-memory operands into two alternating TMEM accumulators; the launch is timed with CUDA events by the caller.
// Operand values are irrelevant (all bytes 1); only the rate matters.  bench.py reports `roofline.peak_int8_measured`
// from it (cnnacc_probe_int8_peak).
#pragma once
#include "conv_fused.cuh"

namespace cnnacc {

constexpr int kProbeSmem = 16384;                     // A: 128 x 32 B, B: 256 x 32 B
constexpr double kProbeOpsPerMma = 2.0 * 128 * 256 * 32;

__global__ void __launch_bounds__(32, 1)
int8_peak_probe_kernel(int iters, int* status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_storage;
    __shared__ uint32_t tmem_slot;
    const uint32_t bar = smem_u32(&bar_storage);
    for (int i = threadIdx.x; i < kProbeSmem / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    tc_fence_after();
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_slot, 0);
    constexpr uint32_t idesc = umma_idesc_i8(256);
    const uint64_t da = umma_desc(smem_u32(smem), 2048, 128);            // packed K-major core matrices
    const uint64_t db = umma_desc(smem_u32(smem) + 4096, 4096, 128);
    for (int i = 0; i < iters; i += 8) {
        if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 8; j++) umma_i8(tm + (uint32_t)(j & 1) * 256u, da, db, idesc, 1u);
        }
        __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    if (!mbar_wait(bar, 0, kWaitBudgetClk) && threadIdx.x == 0) atomicOr(status, 1);
    tc_fence_before();
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
}

}  // namespace cnnacc
