// probe_imma.cu -- throughput of the LEGACY warp-level int8 MMA (mma.sync.m16n8k32.s32.u8.s8) on sm_100a,
// as a candidate for layer 0 (accumulators land in registers, so the dp4a warps could keep their structure).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_imma tools/probe_imma.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void imma_rate(int iters, long long* cycles, int* sink, uint32_t seed) {
    uint32_t a[4] = {seed + threadIdx.x, seed * 3u, seed ^ 0x55u, seed + 7u};
    uint32_t b[8][2];
    int c[8][4];
    for (int j = 0; j < 8; j++) { b[j][0] = seed * (j + 1); b[j][1] = seed + j; for (int i = 0; i < 4; i++) c[j][i] = 0; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) imma(c[j], a, b[j]);
    }
    __syncthreads();
    long long t1 = clock64();
    int s = 0;
    for (int j = 0; j < 8; j++) for (int i = 0; i < 4; i++) s += c[j][i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    long long* dCyc; int* dSink;
    CK(cudaMalloc(&dCyc, 8)); CK(cudaMalloc(&dSink, 148 * 1024 * 4));
    for (int threads : {32, 128, 256, 512, 1024}) {
        const int iters = 4000;
        imma_rate<<<1, threads>>>(iters, dCyc, dSink, 12345u);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost));
        const double n = (double)iters * 8 * (threads / 32);
        printf("mma.sync m16n8k32 u8*s8, %2d warps on one SM: %.2f cycles per IMMA per SM, %.0f MAC/clk/SM\n", threads / 32, c / n, n * 4096 / c);
    }
    // all SMs busy (clock/power effects)
    imma_rate<<<148, 512>>>(4000, dCyc, dSink, 777u);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dCyc, 8, cudaMemcpyDeviceToHost));
    printf("mma.sync m16n8k32, 148 CTAs x 16 warps: %.0f MAC/clk/SM (block 0)\n", 4000.0 * 8 * 16 * 4096 / c);
    return 0;
}
