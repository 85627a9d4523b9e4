// tail.cuh -- follow-on kernel: 4x4 spatial-bin pool -> linear -> softmax -> argmax -> CAM bbox.
//
// One CTA (256 threads) per image, reading the 64x16x16 u8 feature map (16 KiB) once from HBM/L2.
// Reference (all /root/reference/software/realtime_detect.py):
//   classify_vec :68-82   pooled[ch*16 + r*4 + c] = mean of the 4x4 bin of (feat/255); logits = W.pooled + b;
//                         softmax; argmax
//   bbox_vec     :85-116  channels with mean > 250 masked; cam = sum_ch w[cls][ch,bin] * fm (fp32, channel
//                         order); ReLU; /max; thr = max(percentile70, 0.25); bbox of cam > thr, x8 scaling
// Numerics:
//   * bin sums are exact integers; pooled = S / 4080 in one rounding (identical to Classifier.classify's
//     mean-then-/255, pynq_inference.py:325-334; within 1 ulp of classify_vec's /255-then-mean);
//   * logits: fp32, fixed summation tree (4 bins per thread, warp shuffle tree, 8 warp partials in order);
//   * CAM: products and sums rounded separately (no FMA) in channel order 0..63, as numpy's reduction over
//     the outer axis does, so the bbox integers match bit for bit given the same class;
//   * percentile(70) of 256 values = index 178.5 -> hi - (hi-lo)*0.5 in fp32 (numpy _lerp with t = 0.5).
#pragma once
#include "common.cuh"

namespace cnnacc {

constexpr int kMaxClasses = 16;

__global__ void __launch_bounds__(256)
classify_bbox_kernel(const uint8_t* __restrict__ feats, const float* __restrict__ fc_w,
                     const float* __restrict__ fc_b, int n_cls,
                     float* __restrict__ probs, int32_t* __restrict__ cls_out, int32_t* __restrict__ bbox_out,
                     const int32_t* __restrict__ cls_in)
{
    __shared__ __align__(16) uint8_t s_feat[64 * 256];
    __shared__ float s_part[8][kMaxClasses];
    __shared__ float s_cam[256];
    __shared__ __align__(16) float s_wc[1024];           // class weights of the chosen class, masked
    __shared__ float s_red[8];
    __shared__ int   s_valid[64];
    __shared__ int   s_cls;
    __shared__ float s_lohi[2];
    __shared__ int   s_box[4];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t img = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(feats + img * 16384);

    // thread t owns channel t/4, bin-row t%4: four 16-byte map rows = 64 contiguous bytes
    int S[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint4 v = src[t * 4 + r];
        reinterpret_cast<uint4*>(s_feat)[t * 4 + r] = v;
        S[0] = __dp4a(v.x, 0x01010101u, (unsigned)S[0]);
        S[1] = __dp4a(v.y, 0x01010101u, (unsigned)S[1]);
        S[2] = __dp4a(v.z, 0x01010101u, (unsigned)S[2]);
        S[3] = __dp4a(v.w, 0x01010101u, (unsigned)S[3]);
    }
    // channel mean <= 250  <=>  channel sum <= 64000 (bbox_vec: valid = ch_means <= 250)
    int chsum = S[0] + S[1] + S[2] + S[3];
    chsum += __shfl_xor_sync(0xffffffffu, chsum, 1);
    chsum += __shfl_xor_sync(0xffffffffu, chsum, 2);
    if ((t & 3) == 0) s_valid[t >> 2] = (chsum <= 250 * 256);

    float pooled[4];
#pragma unroll
    for (int c = 0; c < 4; c++) pooled[c] = __fdiv_rn((float)S[c], 4080.0f);

    // logits: W row-major [n_cls][1024]; this thread's bins are 4t .. 4t+3
    for (int k = 0; k < n_cls; k++) {
        float4 w = reinterpret_cast<const float4*>(fc_w + (size_t)k * 1024)[t];
        float p = pooled[0] * w.x;
        p = fmaf(pooled[1], w.y, p);
        p = fmaf(pooled[2], w.z, p);
        p = fmaf(pooled[3], w.w, p);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
        if (lane == 0) s_part[warp][k] = p;
    }
    __syncthreads();

    if (warp == 0) {
        float logit = -INFINITY;
        if (lane < n_cls) {
            float a = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; w8++) a += s_part[w8][lane];
            logit = a + fc_b[lane];
        }
        float mx = logit;
        int arg = lane < n_cls ? lane : 0x7fffffff;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float om = __shfl_xor_sync(0xffffffffu, mx, off);
            int   oa = __shfl_xor_sync(0xffffffffu, arg, off);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }   // first maximum, like np.argmax
        }
        float e = lane < n_cls ? expf(logit - mx) : 0.f;
        float sum = e;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (probs && lane < n_cls) probs[img * n_cls + lane] = __fdiv_rn(e, sum);
        if (lane == 0) {
            // bbox_vec takes the class as an argument (realtime_detect.py:85); cls_in carries it when given
            s_cls = cls_in ? min(max(cls_in[img], 0), n_cls - 1) : arg;
            if (cls_out && !cls_in) cls_out[img] = arg;
            s_box[0] = 16; s_box[1] = 16; s_box[2] = -1; s_box[3] = -1;   // min col, min row, max col, max row
        }
    }
    __syncthreads();
    if (!bbox_out) return;

    // CAM: thread t owns pixel t = (py, px); class weights indexed [ch*16 + (py/4)*4 + px/4].  The masked class weights
    // (0 for saturated channels) are staged once per image; u8 -> f32 goes through the 2^23 mantissa trick (one LOP3 +
    // one FADD, exact for 0..255) instead of the quarter-rate I2F.
    reinterpret_cast<float4*>(s_wc)[t] = s_valid[t >> 2] ? __ldg(reinterpret_cast<const float4*>(fc_w + (size_t)s_cls * 1024) + t)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int py = t >> 4, px = t & 15;
    const float* wc = s_wc + (py >> 2) * 4 + (px >> 2);
    float cam = 0.f;
#pragma unroll 16
    for (int ch = 0; ch < 64; ch++) {
        const float f = __fsub_rn(__uint_as_float(0x4B000000u | (uint32_t)s_feat[ch * 256 + t]), 8388608.0f);
        cam = __fadd_rn(cam, __fmul_rn(wc[ch * 16], f));
    }
    cam = fmaxf(cam, 0.f);

    float m = cam;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = s_red[0];
#pragma unroll
    for (int w8 = 1; w8 < 8; w8++) m = fmaxf(m, s_red[w8]);
    if (m > 0.f) cam = __fdiv_rn(cam, m);

    // 70th percentile of 256 values = index 178.5: sorted[178] and sorted[179].  Bitonic sort, one value per thread:
    // strides < 32 exchange by shuffle, strides >= 32 through shared memory (36 compare-exchange stages instead of
    // the 256-step rank-by-counting loop this replaced).
    float v = cam;
#pragma unroll
    for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            float o;
            if (j >= 32) {
                __syncthreads();                         // previous readers of s_cam are done
                s_cam[t] = v;
                __syncthreads();
                o = s_cam[t ^ j];
            } else {
                o = __shfl_xor_sync(0xffffffffu, v, j);
            }
            const bool up = ((t & k) == 0);              // ascending block
            const bool lower = ((t & j) == 0);           // this thread keeps the smaller one in an ascending block
            v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    if (t == 178) s_lohi[0] = v;
    if (t == 179) s_lohi[1] = v;
    __syncthreads();
    const float lo = s_lohi[0], hi = s_lohi[1];
    float thr = __fsub_rn(hi, __fmul_rn(__fsub_rn(hi, lo), 0.5f));
    thr = (0.25f > thr) ? 0.25f : thr;                      // python max(p, 0.25)
    if (cam > thr) {
        atomicMin(&s_box[0], px); atomicMin(&s_box[1], py);
        atomicMax(&s_box[2], px); atomicMax(&s_box[3], py);
    }
    __syncthreads();
    if (t == 0) {
        int4 b;
        if (s_box[2] >= 0) {
            b.x = s_box[0] * 8; b.y = s_box[1] * 8;
            b.z = min(127, (s_box[2] + 1) * 8); b.w = min(127, (s_box[3] + 1) * 8);
        } else {
            b = make_int4(0, 0, 127, 127);
        }
        reinterpret_cast<int4*>(bbox_out)[img] = b;
    }
}

// Spatial-bin pooled features alone: (64,256) u8 -> (1024,) f32, pooled[ch*16 + r*4 + c] = mean(4x4 bin) / 255 -- the
// classifier's input, which the reference's trainer builds from a feature dump (retrain_classifier.py:188-205) and
// Classifier.classify builds per image (pynq_inference.py:325-334).  Bin sums are exact integers and sum/16 is exact,
// so one correctly rounded division S/4080 equals the reference's mean-then-/255.
__global__ void __launch_bounds__(256)
pool_features_kernel(const uint8_t* __restrict__ feats, float* __restrict__ pooled)
{
    const int t = threadIdx.x;
    const size_t img = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(feats + img * 16384);
    int S[4] = {0, 0, 0, 0};                     // thread t: channel t/4, bin-row t%4 -> bins 4t .. 4t+3
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint4 v = __ldg(src + t * 4 + r);
        S[0] = __dp4a(v.x, 0x01010101u, (unsigned)S[0]);
        S[1] = __dp4a(v.y, 0x01010101u, (unsigned)S[1]);
        S[2] = __dp4a(v.z, 0x01010101u, (unsigned)S[2]);
        S[3] = __dp4a(v.w, 0x01010101u, (unsigned)S[3]);
    }
    reinterpret_cast<float4*>(pooled + img * 1024)[t] =
        make_float4(__fdiv_rn((float)S[0], 4080.0f), __fdiv_rn((float)S[1], 4080.0f),
                    __fdiv_rn((float)S[2], 4080.0f), __fdiv_rn((float)S[3], 4080.0f));
}

}  // namespace cnnacc
