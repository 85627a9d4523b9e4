// probe_pipeline.cu -- where does a chunked H2D -> kernel -> D2H pipeline lose PCIe bandwidth?
// Variants (all: 64 chunks of CHUNK MiB each way, pinned host memory, one H2D stream, one D2H stream, one kernel stream):
//   A  copies only, the two directions independent
//   B  A + an event record after every copy
//   C  B + every D2H waits for the matching H2D's event
//   D  H2D -> tiny kernel -> D2H, chained by events
//   E  D with a kernel that keeps all SMs busy for ~chunk time / 4
//   F  E, but H2D stream also waits for the D2H event of chunk i-4 (slot reuse)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/probe_pipeline tools/probe_pipeline.cu
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void tiny(const uint8_t* in, uint8_t* out) { if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = in[0]; }
__global__ void busy(const uint4* in, uint4* out, size_t n16, int reps) {
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v = in[i]; v.x += r; out[i] = v;
        }
}

int main(int argc, char** argv) {
    const size_t chunk = (size_t)(argc > 1 ? atoi(argv[1]) : 4) << 20;
    const int nchunk = 64, slots = 4;
    uint8_t *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, chunk * nchunk, cudaHostAllocPortable)); CK(cudaHostAlloc(&h_out, chunk * nchunk, cudaHostAllocPortable));
    CK(cudaMalloc(&d_in, chunk * nchunk)); CK(cudaMalloc(&d_out, chunk * nchunk));
    cudaStream_t sh, sk, sd;
    CK(cudaStreamCreateWithFlags(&sh, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ein(nchunk), ek(nchunk), eout(nchunk);
    for (int i = 0; i < nchunk; i++) { CK(cudaEventCreateWithFlags(&ein[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ek[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&eout[i], cudaEventDisableTiming)); }
    for (char v = 'A'; v <= 'F'; v++) {
        double best = 0;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaDeviceSynchronize());
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < nchunk; i++) {
                const size_t o = (size_t)i * chunk;
                if (v == 'F' && i >= slots) CK(cudaStreamWaitEvent(sh, eout[i - slots], 0));
                CK(cudaMemcpyAsync(d_in + o, h_in + o, chunk, cudaMemcpyHostToDevice, sh));
                if (v >= 'B') CK(cudaEventRecord(ein[i], sh));
                if (v >= 'D') {
                    CK(cudaStreamWaitEvent(sk, ein[i], 0));
                    if (v == 'D') tiny<<<1, 32, 0, sk>>>(d_in + o, d_out + o);
                    else busy<<<148 * 4, 256, 0, sk>>>((const uint4*)(d_in + o), (uint4*)(d_out + o), chunk / 16, 8);
                    CK(cudaEventRecord(ek[i], sk));
                    CK(cudaStreamWaitEvent(sd, ek[i], 0));
                } else if (v == 'C') CK(cudaStreamWaitEvent(sd, ein[i], 0));
                CK(cudaMemcpyAsync(h_out + o, d_out + o, chunk, cudaMemcpyDeviceToHost, sd));
                if (v >= 'B') CK(cudaEventRecord(eout[i], sd));
            }
            const double t_issue = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            CK(cudaDeviceSynchronize());
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const double gbs = chunk * nchunk / dt / 1e9;
            if (gbs > best) best = gbs;
            if (rep == 2) printf("variant %c chunk %zu MiB: %.1f GB/s per direction (best of 3), %.0f us per chunk, host issue %.1f us per chunk\n",
                                 v, chunk >> 20, best, chunk / (best * 1e3), t_issue / nchunk * 1e6);
        }
    }
    return 0;
}
