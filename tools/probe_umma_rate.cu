// probe_umma_rate.cu -- where does the ~77-cycle floor of an M=128 kind::i8 MMA from shared memory come from?
// Times back-to-back tcgen05.mma (one issuing thread, two alternating accumulators) for several operand shapes
// and shared-memory layouts.  Values are garbage on purpose; only cycles matter.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_umma_rate tools/probe_umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, long long timeout = 400000000LL) {
    long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (clock64() - t0 > timeout) return false;
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 |
           (uint64_t)1 << 46 | (uint64_t)layout << 61;
}
__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Cfg { int M, N, a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout, a_step, n_acc, other_warps; };

constexpr int kSmem = 200 * 1024;

__global__ void __launch_bounds__(512) rate_kernel(Cfg c, int iters, long long* cycles, int* status, uint32_t* sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < kSmem / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x02020202u, 0, 0x01010101u);
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);       // warp-uniform for the compiler
    if (warp == 0) {
        // the whole warp runs the loop; one elected thread issues.  Uniform operands keep the issue path on the
        // uniform datapath (no per-MMA R2UR waterfall), so the issuing thread is not the limit.
        const uint32_t idesc = make_idesc_i8(c.M, c.N);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 120 * 1024;
        const uint64_t da0 = make_desc(a_base, c.a_lbo, c.a_sbo, c.a_layout);
        const uint64_t db = make_desc(b_base, c.b_lbo, c.b_sbo, c.b_layout);
        const uint64_t astep = (uint64_t)(c.a_step >> 4);
        const uint32_t dstep = (c.n_acc > 1) ? (uint32_t)(c.N < 256 ? 128 : 256) : 0u;
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t d = tm + (uint32_t)(j & 1) * dstep + ((c.n_acc == 4) ? (uint32_t)((j >> 1) & 1) * 256u : 0u);
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                                 :: "r"(d), "l"(da0 + astep * j), "l"(db), "r"(idesc), "r"(1) : "memory");
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        bool ok = mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; if (!ok) *status = 2; stop = 1; }
    } else if (threadIdx.x >= 32 && (int)threadIdx.x < 32 + 32 * c.other_warps) {
        // competing shared-memory traffic: what the dp4a warps' LDS/STS would do to the operand fetch
        uint32_t acc = 0;
        const uint4* p = reinterpret_cast<const uint4*>(smem + 160 * 1024);
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; j++) { uint4 v = p[(threadIdx.x + 64 * j) & 1023]; acc += v.x + v.y + v.z + v.w; }
        }
        sink[threadIdx.x] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
}

int main() {
    long long* dCyc; int* dStatus; uint32_t* dSink;
    CK(cudaMalloc(&dCyc, 8 * 256)); CK(cudaMalloc(&dStatus, 4)); CK(cudaMalloc(&dSink, 4096));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    struct Named { const char* name; Cfg c; };
    const Named tests[] = {
        // name                                   M    N   a_lbo  a_sbo lay  b_lbo b_sbo lay a_step acc other
        {"packed A, N=32",                      {128,  32, 2048,   128, 0,   512, 128, 0,    0, 2, 0}},
        {"packed A, N=64",                      {128,  64, 2048,   128, 0,  1024, 128, 0,    0, 2, 0}},
        {"packed A, N=128",                     {128, 128, 2048,   128, 0,  2048, 128, 0,    0, 2, 0}},
        {"packed A, N=256",                     {128, 256, 2048,   128, 0,  4096, 128, 0,    0, 2, 0}},
        {"packed A, N=128, M=64",               { 64, 128, 1024,   128, 0,  2048, 128, 0,    0, 2, 0}},
        {"packed A, N=32,  M=64",               { 64,  32, 1024,   128, 0,   512, 128, 0,    0, 2, 0}},
        {"L1 Toeplitz A (lbo 528, sbo 2112) N=128", {128, 128, 528, 2112, 0, 2048, 128, 0, 1056, 2, 0}},
        {"L2 tap A (lbo 18496, sbo 1088) N=64", {128,  64, 18496, 1088, 0,  1024, 128, 0,   16, 2, 0}},
        {"L2 tap A, N=64, 4 accumulators",      {128,  64, 18496, 1088, 0,  1024, 128, 0,   16, 4, 0}},
        {"L1 Toeplitz A, N=128, 1 accumulator", {128, 128, 528, 2112, 0,  2048, 128, 0, 1056, 1, 0}},
        {"swizzle-128B A+B (sbo 1024), N=128",  {128, 128, 16,   1024, 2,    16, 1024, 2,   32, 2, 0}},
        {"swizzle-128B A+B (sbo 1024), N=64",   {128,  64, 16,   1024, 2,    16, 1024, 2,   32, 2, 0}},
        {"swizzle-128B A+B (sbo 1024), N=256",  {128, 256, 16,   1024, 2,    16, 1024, 2,   32, 2, 0}},
        {"swizzle-32B A (sbo 256), N=128",      {128, 128, 16,    256, 6,  2048, 128, 0,    0, 2, 0}},
        {"L1 Toeplitz A, N=128 + 8 LDS warps",  {128, 128, 528, 2112, 0,  2048, 128, 0, 1056, 2, 8}},
        {"L1 Toeplitz A, N=128 + 15 LDS warps", {128, 128, 528, 2112, 0,  2048, 128, 0, 1056, 2, 15}},
    };
    for (int grid : {1, 148}) {
        printf("---- %d CTA(s) (one per SM) ----\n", grid);
        for (const Named& t : tests) {
            const int iters = 4000;
            CK(cudaMemset(dStatus, 0, 4));
            rate_kernel<<<grid, 512, kSmem>>>(t.c, iters, dCyc, dStatus, dSink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", t.name, cudaGetErrorString(e)); return 1; }
            long long c[148]; int st;
            CK(cudaMemcpy(c, dCyc, 8 * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dStatus, 4, cudaMemcpyDeviceToHost));
            long long mx = 0, mn = 1LL << 60;
            for (int i = 0; i < grid; i++) { mx = c[i] > mx ? c[i] : mx; mn = c[i] < mn ? c[i] : mn; }
            printf("%-44s : %6.1f cycles/MMA (min over CTAs %6.1f)  (%5.0f MAC/clk/SM)%s\n", t.name, (double)mx / iters, (double)mn / iters,
                   (double)t.c.M * t.c.N * 32 * iters / mx, st ? "  [TIMEOUT]" : "");
        }
    }
    return 0;
}
