// probe_tma.cu -- which 3D tiled TMA box forms load a zero-padded 128x128 u8 image correctly on sm_100a?
// usage: probe_tma <variant>   (one variant per process: a fault poisons the context)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

struct V { int rank; int bx, by, cx, cy; int use_tile; int smem_off; };

__global__ void k(const __grid_constant__ CUtensorMap map, V v, int img, uint8_t* out, int* status)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem) + v.smem_off;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    const int bytes = v.bx * v.by;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) smem[i] = 0xEE;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
        if (v.rank == 3) {
            if (v.use_tile)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(sb), "l"(&map), "r"(v.cx), "r"(v.cy), "r"(img), "r"(b) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             :: "r"(sb), "l"(&map), "r"(v.cx), "r"(v.cy), "r"(img), "r"(b) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(sb), "l"(&map), "r"(v.cx), "r"(v.cy + img * 128), "r"(b) : "memory");
        }
    }
    long long t0 = clock64(); bool ok = false;
    while (clock64() - t0 < 100000000LL) {
        uint32_t p;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(p) : "r"(b) : "memory");
        if (p) { ok = true; break; }
    }
    if (!ok && threadIdx.x == 0) *status = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[v.smem_off + i];
}

int main(int argc, char** argv)
{
    int var = argc > 1 ? atoi(argv[1]) : 0;
    static const V vars[] = {
        {3, 144, 130, -4, -1, 1, 0},      // 0: the fused kernel's form
        {3, 128, 128, 0, 0, 1, 0},        // 1: exact box, no OOB
        {3, 144, 130, 0, 0, 1, 0},        // 2: oversize box, OOB on the high side only
        {3, 160, 130, -16, -1, 1, 0},     // 3: 16-byte aligned negative x
        {3, 144, 130, -4, -1, 0, 0},      // 4: no .tile qualifier
        {2, 144, 130, -4, -1, 0, 0},      // 5: 2D map over [n*128][128]
        {3, 144, 128, -4, 0, 1, 0},       // 6: negative x only
        {3, 128, 130, 0, -1, 1, 0},       // 7: negative y only
        {3, 144, 130, -4, -1, 1, 18816},  // 8: form 0 into the second slot
        {3, 144, 65, -4, -1, 1, 0},       // 9: half-height box
    };
    V v = vars[var];
    const int n = 4;
    std::vector<uint8_t> h(n * 16384);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(1 + (i * 7 + i / 128) % 250);
    uint8_t *d, *dout; int* dst;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&dout, 65536)); CK(cudaMalloc(&dst, 4));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice)); CK(cudaMemset(dst, 0, 4)); CK(cudaMemset(dout, 0xCC, 65536));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    auto enc = (CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fp;
    CUtensorMap map;
    CUresult r;
    if (v.rank == 3) {
        cuuint64_t gd[3] = {128, 128, (cuuint64_t)n}; cuuint64_t gs[2] = {128, 16384};
        cuuint32_t box[3] = {(cuuint32_t)v.bx, (cuuint32_t)v.by, 1}; cuuint32_t es[3] = {1, 1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t gd[2] = {128, (cuuint64_t)n * 128}; cuuint64_t gs[1] = {128};
        cuuint32_t box[2] = {(cuuint32_t)v.bx, (cuuint32_t)v.by}; cuuint32_t es[2] = {1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    printf("variant %d: rank %d box %dx%d at (%d,%d) tile=%d smem_off=%d : encode=%d ", var, v.rank, v.bx, v.by, v.cx, v.cy, v.use_tile, v.smem_off, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 2; }
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int img = 2;
    k<<<1, 128, 65536>>>(map, v, img, dout, dst);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("KERNEL FAULT: %s\n", cudaGetErrorString(e)); return 3; }
    int st; CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> o(v.bx * v.by);
    CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int y = 0; y < v.by; y++)
        for (int x = 0; x < v.bx; x++) {
            int gx = v.cx + x, gy = v.cy + y;
            uint8_t want = 0;
            if (v.rank == 3) { if (gx >= 0 && gx < 128 && gy >= 0 && gy < 128) want = h[img * 16384 + gy * 128 + gx]; }
            else { int gy2 = gy + img * 128; if (gx >= 0 && gx < 128 && gy2 >= 0 && gy2 < n * 128) want = h[gy2 * 128 + gx]; }
            bad += o[y * v.bx + x] != want;
        }
    printf("%s%s mismatches=%ld\n", st ? "TIMEOUT " : "", bad ? "FAIL" : "PASS", bad);
    return 0;
}
