// cam_upsampled.cuh -- the other CAM box of the reference: Classifier.get_cam_bbox
// (/root/reference/software/pynq_inference.py:349-408), batched, one CTA (256 threads) per image.
//
//   feat (64,16,16) u8, class index, class weights (1024,) f32
//     -> channels with mean > 250 skipped; cam = sum_ch w[ch,bin] * fm (fp32, channel order, product and sum rounded
//        separately); ReLU; / max                                                   (:355-381)
//     -> (cam * 255).astype(uint8)                                                   (:384)
//     -> PIL resize 16x16 -> 128x128, BILINEAR, mode "L"                             (:385)
//     -> threshold = max(percentile70(cam_u8 / 255), 0.2); box of cam > threshold, padded by 3, clipped   (:389-404)
//
// Pillow's 8-bit resampler is integer arithmetic (src/libImaging/Resample.c): per output index a window [xmin, xmin+xmax)
// and coefficients rounded to 22 fractional bits; out = clip8((2^21 + sum in*k) >> 22); horizontal pass first, rounded to
// u8, then vertical.  The coefficient table is built on the host in double precision exactly as precompute_coeffs /
// normalize_coeffs_8bpc do (make_resample_tab below) and handed to the kernel as a parameter.
//
// The threshold needs no floating point: cam_full = level/255 takes 256 values, np.percentile interpolates between the
// sorted elements 11468 and 11469 (0.7 * 16383 = 11468.1) and so lands in [k_lo/255, k_hi/255), and 51/255 == 0.2f; hence
// mask = level > max(k_lo, 51) with k_lo = the level of sorted[11468] (oracle/np_oracle.py get_cam_bbox_levels, checked
// against the float formulation in tests/test_oracle.py).  k_lo comes from an 8-step bisection on "how many levels <= mid",
// each thread counting over the 64 output pixels it keeps in registers (four per SIMD-in-word compare).
#pragma once
#include <cmath>
#include "common.cuh"

namespace cnnacc {

constexpr int kCamIn = 16, kCamOut = 128, kCamTaps = 3, kPilBits = 22;

struct ResampleTab {
    int32_t k[kCamOut][kCamTaps];      // fixed-point coefficients, zero beyond the window
    int8_t  xmin[kCamOut];
};

// Pillow: precompute_coeffs(inSize=16, in0=0, in1=16, outSize=128, bilinear) + normalize_coeffs_8bpc.
inline ResampleTab make_resample_tab() {
    ResampleTab tab{};
    const double scale = (double)kCamIn / kCamOut;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    for (int xx = 0; xx < kCamOut; xx++) {
        const double center = 0.0 + (xx + 0.5) * scale;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > kCamIn) xmax = kCamIn;
        xmax -= xmin;
        double w[kCamTaps] = {0.0, 0.0, 0.0}, ww = 0.0;
        for (int x = 0; x < xmax; x++) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w[x] = a < 1.0 ? 1.0 - a : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; x++)
            if (ww != 0.0) w[x] /= ww;
        for (int x = 0; x < kCamTaps; x++)
            tab.k[xx][x] = w[x] < 0 ? (int)(-0.5 + w[x] * (1 << kPilBits)) : (int)(0.5 + w[x] * (1 << kPilBits));
        tab.xmin[xx] = (int8_t)xmin;
    }
    return tab;
}

__device__ __forceinline__ uint32_t pil_clip8(int acc) { return (uint32_t)min(max(acc >> kPilBits, 0), 255); }

__global__ void __launch_bounds__(256)
cam_bbox_upsampled_kernel(const uint8_t* __restrict__ feats, const float* __restrict__ fc_w, int n_cls,
                          const int32_t* __restrict__ cls_in, int32_t* __restrict__ bbox_out, uint8_t* __restrict__ cam_out,
                          const __grid_constant__ ResampleTab tab)
{
    __shared__ __align__(16) uint8_t s_feat[64 * 256];
    __shared__ __align__(16) float s_wc[1024];
    __shared__ float s_red[8];
    __shared__ int   s_valid[64];
    __shared__ uint8_t s_q[kCamIn][kCamIn];              // (cam * 255).astype(uint8)
    __shared__ __align__(4) uint8_t s_h[kCamIn][kCamOut]; // after the horizontal pass
    __shared__ int   s_cnt[8];
    __shared__ int   s_box[4];

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t img = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(feats + img * 16384);

    // thread t owns channel t/4, rows 4*(t%4) .. +3 (64 contiguous bytes); channel sum for the saturation test
    int chsum = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint4 v = src[t * 4 + r];
        reinterpret_cast<uint4*>(s_feat)[t * 4 + r] = v;
        chsum = __dp4a(v.x, 0x01010101u, (unsigned)chsum); chsum = __dp4a(v.y, 0x01010101u, (unsigned)chsum);
        chsum = __dp4a(v.z, 0x01010101u, (unsigned)chsum); chsum = __dp4a(v.w, 0x01010101u, (unsigned)chsum);
    }
    chsum += __shfl_xor_sync(0xffffffffu, chsum, 1);
    chsum += __shfl_xor_sync(0xffffffffu, chsum, 2);
    if ((t & 3) == 0) s_valid[t >> 2] = (chsum <= 250 * 256);        // ch_means[ch] > 250 -> skipped (:367-368)
    if (t == 0) { s_box[0] = kCamOut; s_box[1] = kCamOut; s_box[2] = -1; s_box[3] = -1; }
    __syncthreads();

    const int cls = min(max(cls_in[img], 0), n_cls - 1);
    reinterpret_cast<float4*>(s_wc)[t] = s_valid[t >> 2] ? __ldg(reinterpret_cast<const float4*>(fc_w + (size_t)cls * 1024) + t)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // CAM, thread t = pixel (py, px); a skipped channel adds +0.0, which leaves the running sum unchanged
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
    if (m > 0.f) cam = div_rn_zero_ok(cam, m);
    s_q[py][px] = (uint8_t)(int)__fmul_rn(cam, 255.0f);               // astype(uint8): truncation, value in 0..255
    __syncthreads();

    // horizontal pass: 16 rows x 128 columns, 8 values per thread
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int idx = t * 8 + j, row = idx >> 7, xx = idx & 127, x0 = tab.xmin[xx];
        int acc = 1 << (kPilBits - 1);
#pragma unroll
        for (int x = 0; x < kCamTaps; x++) acc += (int)s_q[row][min(x0 + x, kCamIn - 1)] * tab.k[xx][x];   // k = 0 beyond the window
        s_h[row][xx] = (uint8_t)pil_clip8(acc);
    }
    __syncthreads();

    // vertical pass: thread owns columns 4*(t%32) .. +3 of rows t/32 + 8*i; 16 packed words stay in registers
    const int xq = lane * 4;
    uint32_t up[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int yy = warp + 8 * i, y0 = tab.xmin[yy];
        int a0, a1, a2, a3;
        a0 = a1 = a2 = a3 = 1 << (kPilBits - 1);
#pragma unroll
        for (int y = 0; y < kCamTaps; y++) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(&s_h[min(y0 + y, kCamIn - 1)][xq]);
            const int k = tab.k[yy][y];
            a0 += (int)(v & 255u) * k; a1 += (int)((v >> 8) & 255u) * k;
            a2 += (int)((v >> 16) & 255u) * k; a3 += (int)(v >> 24) * k;
        }
        up[i] = pil_clip8(a0) | (pil_clip8(a1) << 8) | (pil_clip8(a2) << 16) | (pil_clip8(a3) << 24);
        if (cam_out) *reinterpret_cast<uint32_t*>(cam_out + img * (size_t)(kCamOut * kCamOut) + yy * kCamOut + xq) = up[i];
    }
    if (!bbox_out) return;

    // k_lo = level of sorted[11468]: the smallest level L with #(values <= L) >= 11469
    constexpr int kRank = 11469;          // floor(0.7 * (128*128 - 1)) + 1
    int lo = 0, hi = 255;
#pragma unroll 1
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        // four pixels per compare: __vsetgtu4 leaves 1 in every byte lane whose pixel is > mid; the lanes add up over the 16
        // words without carrying (<= 16 each) and dp4a sums them
        const uint32_t mid4 = (uint32_t)mid * 0x01010101u;
        uint32_t gt4 = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) gt4 += __vsetgtu4(up[i], mid4);
        int c = 64 - (int)__dp4a(gt4, 0x01010101u, 0u);       // pixels <= mid among this thread's 64
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        __syncthreads();                                   // previous round's readers of s_cnt are done
        if (lane == 0) s_cnt[warp] = c;
        __syncthreads();
        c = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; w8++) c += s_cnt[w8];
        if (c >= kRank) hi = mid; else lo = mid + 1;
    }
    const int thr = max(lo, 51);                           // max(threshold, 0.2): 51/255 == 0.2f

    // box of the pixels > thr: per word a byte-lane mask (0xFF where the pixel is above); the OR of the masks gives the thread's
    // columns, the first / last non-empty word its rows (yy grows with i)
    int bx0 = kCamOut, by0 = kCamOut, bx1 = -1, by1 = -1;
    const uint32_t thr4 = (uint32_t)thr * 0x01010101u;
    uint32_t cols = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t m4 = __vcmpgtu4(up[i], thr4);
        const int yy = warp + 8 * i;
        cols |= m4;
        if (m4) { by0 = min(by0, yy); by1 = yy; }
    }
    if (cols) { bx0 = xq + ((__ffs((int)cols) - 1) >> 3); bx1 = xq + ((31 - __clz((int)cols)) >> 3); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, off)); by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, off));
        bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, off)); by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, off));
    }
    if (lane == 0 && bx1 >= 0) {
        atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1);
    }
    __syncthreads();
    if (t == 0) {
        int4 b;
        if (s_box[2] >= 0) {
            b.x = max(0, s_box[0] - 3); b.y = max(0, s_box[1] - 3);
            b.z = min(kCamOut - 1, s_box[2] + 3); b.w = min(kCamOut - 1, s_box[3] + 3);
        } else {
            b = make_int4(0, 0, kCamOut - 1, kCamOut - 1);
        }
        reinterpret_cast<int4*>(bbox_out)[img] = b;
    }
}

}  // namespace cnnacc
