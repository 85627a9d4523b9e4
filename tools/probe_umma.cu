// probe_umma.cu -- one-shot hardware probe for the fused kernel's building blocks (run on the B200 box).
//
//   1. correctness of tcgen05.mma kind::i8 (A = u8, B = s8, D = s32 in TMEM) with hand-built no-swizzle
//      K-major shared-memory descriptors, including 16-byte-granular start offsets (conv tap shifts),
//      LBO used as an arbitrary "second K half" offset (tap pairing) and SBO = 2 x row pitch;
//   2. issue/throughput of M=128 MMAs with N = 32 / 64 / 128 / 256 from shared memory;
//   3. tcgen05.ld throughput;   4. dp4a (IDP.4A) throughput;
// Every wait has a cycle-count timeout so a wrong descriptor cannot hang the GPU.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_umma tools/probe_umma.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, long long timeout = 50000000LL) {
    long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (clock64() - t0 > timeout) return false;
    }
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version 1 (sm_100)
    return d;                               // layout type 0 = no swizzle, base offset 0
}
__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N) {
    // c_format S32 = 2 @ [4,6); a_format u8 = 0 @ [7,10); b_format s8 = 1 @ [10,13); K-major both; N>>3 @ [17,23); M>>4 @ [24,29)
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

struct MmaOp { uint32_t a_off, b_off; };
struct Test {
    uint32_t lbo_a, sbo_a, lbo_b, sbo_b;
    int N, n_ops;
    MmaOp ops[16];
};

constexpr int kABytes = 96 * 1024, kBBytes = 32 * 1024;

// ---------------- correctness kernel: 128 threads, one CTA -------------------------------------------
__global__ void __launch_bounds__(128) umma_check_kernel(const uint8_t* a_img, const uint8_t* b_img, Test t, int32_t* d_out, int* status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + kABytes;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;

    for (int i = threadIdx.x; i < kABytes / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < kBBytes / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;

    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_i8(128, t.N);
        for (int i = 0; i < t.n_ops; i++) {
            uint64_t da = make_desc(smem_u32(sA) + t.ops[i].a_off, t.lbo_a, t.sbo_a);
            uint64_t db = make_desc(smem_u32(sB) + t.ops[i].b_off, t.lbo_b, t.sbo_b);
            umma_i8(tm, da, db, idesc, i > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    bool ok = mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    if (!ok) { if (threadIdx.x == 0) *status = 1; }
    else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c0 = 0; c0 < t.N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; j++) d_out[(warp * 32 + lane) * 256 + c0 + j] = (int32_t)v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(256u) : "memory");
}

// ---------------- MMA throughput: one issuing thread, `iters` MMAs, alternating 2 accumulators ---------
__global__ void __launch_bounds__(128) umma_rate_kernel(int N, int iters, int a_stride, long long* cycles, int* status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (kABytes + kBBytes) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x02020202u, 0, 0x01010101u);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_i8(128, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            uint32_t aoff = (uint32_t)((i * a_stride) % (kABytes - 20 * 1024)) & ~15u;
            uint64_t da = make_desc(smem_u32(smem) + aoff, 1056 * 16, 1056 * 2);
            uint64_t db = make_desc(smem_u32(smem) + kABytes, 1024, 128);
            umma_i8(tm + (i & 1) * 256, da, db, idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        bool ok = mbar_wait(smem_u32(&bar), 0, 400000000LL);
        long long t1 = clock64();
        cycles[0] = t1 - t0;
        if (!ok) *status = 2;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
}

// ---------------- tcgen05.ld throughput: nwarps warps each read `iters` x (32 lanes x 32 columns) -------
__global__ void tmem_ld_rate_kernel(int iters, long long* cycles, uint32_t* sink)
{
    __shared__ uint32_t tmem_base;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    const int warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + ((i * 32) & 511), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) acc ^= v[j];
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
}

// ---------------- dp4a throughput -------------------------------------------------------------------------
__global__ void dp4a_rate_kernel(int iters, long long* cycles, int* sink, uint32_t seed)
{
    int acc[16];
    uint32_t a = seed + threadIdx.x, b = seed * 7 + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++)
            asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a + j), "r"(b));
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    int s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += acc[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------- host ---------------------------------------------------------------------------------------
static int run_test(const char* name, const Test& t, const std::vector<uint8_t>& A, const std::vector<uint8_t>& B,
                    uint8_t* dA, uint8_t* dB, int32_t* dD, int* dStatus)
{
    CK(cudaMemset(dD, 0xFF, 128 * 256 * sizeof(int32_t)));
    CK(cudaMemset(dStatus, 0, sizeof(int)));
    umma_check_kernel<<<1, 128, kABytes + kBBytes>>>(dA, dB, t, dD, dStatus);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", name, cudaGetErrorString(e)); return 2; }
    int st; CK(cudaMemcpy(&st, dStatus, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int32_t> D(128 * 256);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (st) { printf("[%s] TIMEOUT waiting for MMA commit\n", name); return 1; }
    long bad = 0; int first_r = -1, first_n = -1; int32_t got0 = 0, want0 = 0;
    for (int r = 0; r < 128; r++)
        for (int n = 0; n < t.N; n++) {
            int32_t want = 0;
            for (int i = 0; i < t.n_ops; i++)
                for (int k = 0; k < 32; k++) {
                    uint32_t aa = t.ops[i].a_off + (r / 8) * t.sbo_a + (r % 8) * 16 + (k / 16) * t.lbo_a + (k % 16);
                    uint32_t bb = t.ops[i].b_off + (n / 8) * t.sbo_b + (n % 8) * 16 + (k / 16) * t.lbo_b + (k % 16);
                    want += (int32_t)A[aa] * (int32_t)(int8_t)B[bb];
                }
            int32_t got = D[r * 256 + n];
            if (got != want) { if (!bad) { first_r = r; first_n = n; got0 = got; want0 = want; } bad++; }
        }
    printf("[%s] N=%d ops=%d lboA=%u sboA=%u lboB=%u sboB=%u : %s (%ld / %d mismatches", name, t.N, t.n_ops, t.lbo_a, t.sbo_a,
           t.lbo_b, t.sbo_b, bad ? "FAIL" : "PASS", bad, 128 * t.N);
    if (bad) printf("; first at r=%d n=%d got %d want %d", first_r, first_n, got0, want0);
    printf(")\n");
    return bad ? 1 : 0;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
    CK(cudaFuncSetAttribute(umma_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kABytes + kBBytes));
    CK(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kABytes + kBBytes));

    std::vector<uint8_t> A(kABytes), B(kBBytes);
    srand(1);
    for (auto& v : A) v = rand() & 0xFF;
    for (auto& v : B) v = rand() & 0xFF;
    uint8_t *dA, *dB; int32_t* dD; int* dStatus; long long* dCyc; uint32_t* dSink;
    CK(cudaMalloc(&dA, kABytes)); CK(cudaMalloc(&dB, kBBytes)); CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMalloc(&dStatus, 4)); CK(cudaMalloc(&dCyc, 8 * 1024)); CK(cudaMalloc(&dSink, 4 * 1024 * 1024));
    CK(cudaMemcpy(dA, A.data(), kABytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), kBBytes, cudaMemcpyHostToDevice));

    int fails = 0;
    {   // T1: plain packed tile: A 128 rows, K=32: core matrices contiguous (SBO 128 between row groups, LBO 2048 between K halves)
        Test t{}; t.lbo_a = 2048; t.sbo_a = 128; t.lbo_b = 1024; t.sbo_b = 128; t.N = 64; t.n_ops = 1; t.ops[0] = {0, 0};
        fails += run_test("T1 packed", t, A, B, dA, dB, dD, dStatus);
    }
    {   // T2: same, N=32, start offsets that are 16-byte but not 128-byte aligned
        Test t{}; t.lbo_a = 2048; t.sbo_a = 128; t.lbo_b = 512; t.sbo_b = 128; t.N = 32; t.n_ops = 1; t.ops[0] = {16 * 5, 16 * 3};
        fails += run_test("T2 unaligned start", t, A, B, dA, dB, dD, dStatus);
    }
    {   // T3: layer-2 style: act map [icb][34][2][17][16B]: row pitch 544, SBO = 2 rows = 1088, LBO = channel-block plane 18496; 9 taps accumulate
        Test t{}; t.lbo_a = 34 * 544; t.sbo_a = 1088; t.lbo_b = 1024; t.sbo_b = 128; t.N = 64; t.n_ops = 9;
        for (int tap = 0; tap < 9; tap++) {
            int dy = tap / 3, dx = tap % 3;
            t.ops[tap] = {(uint32_t)(dy * 544 + (dx & 1) * 272 + (dx >> 1) * 16), (uint32_t)(tap * 2048)};
        }
        fails += run_test("T3 layer2 taps", t, A, B, dA, dB, dD, dStatus);
    }
    {   // T4: layer-1 style: act map [66][2][33][16B]: row pitch 1056, SBO = 2112, LBO = offset to the paired tap (16 B .. a row)
        Test t{}; t.lbo_a = 528; t.sbo_a = 2112; t.lbo_b = 256; t.sbo_b = 128; t.N = 32; t.n_ops = 5;
        for (int i = 0; i < 5; i++) t.ops[i] = {(uint32_t)(i * 1056 + (i & 1) * 16), (uint32_t)(i * 1024)};
        fails += run_test("T4 layer1 pairs lbo=528", t, A, B, dA, dB, dD, dStatus);
        t.lbo_a = 16;
        fails += run_test("T4b lbo=16", t, A, B, dA, dB, dD, dStatus);
        t.lbo_a = 1056 + 512;
        fails += run_test("T4c lbo=row+512", t, A, B, dA, dB, dD, dStatus);
        t.lbo_a = 0;
        fails += run_test("T4d lbo=0", t, A, B, dA, dB, dD, dStatus);
    }
    {   // T5: N = 128 and 256 with packed B
        Test t{}; t.lbo_a = 2048; t.sbo_a = 128; t.lbo_b = 4096; t.sbo_b = 128; t.N = 128; t.n_ops = 2; t.ops[0] = {0, 0}; t.ops[1] = {4096, 8192};
        fails += run_test("T5 N=128", t, A, B, dA, dB, dD, dStatus);
        t.N = 256; t.lbo_b = 4096; t.n_ops = 1;
        fails += run_test("T5b N=256", t, A, B, dA, dB, dD, dStatus);
    }
    printf("correctness: %d failing tests\n", fails);

    // ---- MMA rate ------------------------------------------------------------------------------------
    for (int N : {32, 64, 128, 256}) {
        for (int stride : {0, 2112}) {
            const int iters = 4000;
            CK(cudaMemset(dStatus, 0, 4));
            umma_rate_kernel<<<1, 128, kABytes + kBBytes>>>(N, iters, stride, dCyc, dStatus);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
            long long c; int st; CK(cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dStatus, 4, cudaMemcpyDeviceToHost));
            printf("umma i8 M=128 N=%3d K=32 a_stride=%4d : %.1f cycles/MMA  (%.0f MAC/clk/SM)%s\n", N, stride, (double)c / iters,
                   128.0 * N * 32 * iters / c, st ? "  [TIMEOUT]" : "");
        }
    }
    // same with all SMs busy (power / clock effect is not visible in cycles, but smem contention per SM is the same) -- skipped

    // ---- tcgen05.ld rate ---------------------------------------------------------------------------------
    for (int threads : {128, 256, 512}) {
        const int iters = 2000;
        tmem_ld_rate_kernel<<<1, threads>>>(iters, dCyc, dSink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tmem ld kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long c; CK(cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost));
        double bytes = (double)iters * (threads / 32) * 32 * 32 * 4;
        printf("tcgen05.ld 32x32b.x32, %2d warps: %.1f cycles per warp-load, %.1f B/clk/SM\n", threads / 32, (double)c / iters, bytes / c);
    }
    // ---- dp4a rate -----------------------------------------------------------------------------------------
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 2000;
        dp4a_rate_kernel<<<1, threads>>>(iters, dCyc, (int*)dSink, 12345u);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("dp4a kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long c; CK(cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost));
        printf("dp4a.u32.s32, %2d warps: %.2f lane-dp4a/clk/SM\n", threads / 32, (double)iters * 16 * threads / c);
    }
    return fails ? 3 : 0;
}
