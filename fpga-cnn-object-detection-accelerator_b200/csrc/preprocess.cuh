// preprocess.cuh -- the step in front of the hot path in the real-time loop
// (/root/reference/software/realtime_detect.py:582-591), batched and fused into one kernel:
//
//     centre-crop the (h,w,3) BGR frame to a square of side S = min(h,w)
//  -> cv2.cvtColor(COLOR_BGR2GRAY)      gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15   (OpenCV's 15-bit fixed point)
//  -> cv2.resize((128,128), INTER_AREA)
//
// OpenCV's INTER_AREA has three arithmetic paths, all reproduced bit for bit (pinned against cv2 4.13.0,
// tests/golden/prep_cases.npz):
//   * S == 128           : copy
//   * S == 256           : ResizeAreaFastVec, (a + b + c + d + 2) >> 2
//   * S == 128 k, k >= 3 : ResizeAreaFast, saturate_cast<uchar>(float(box sum) * (1.f / (k*k)))  (round half to even)
//   * otherwise          : ResizeArea, float accumulation: per source row buf = sum_k S[sx_k] * alpha_k (left to right),
//                          then out = beta_0 * buf_0 + beta_1 * buf_1 + ... (top to bottom), every product and sum rounded
//                          to fp32 on its own (no FMA), weights from computeResizeAreaTab in double -> float.
// The weight table (same for x and y: the crop is square) is built on the host by make_area_tab below.
//
// Work split: one CTA of 128 threads per (output row, frame); thread = output column.  A VGA frame is 900 KiB in and
// 16 KiB out, so the kernel is bound by HBM reads of the frames (and, end to end, by PCIe bringing them in).
#pragma once
#include <cmath>
#include <vector>
#include "common.cuh"

namespace cnnacc {

constexpr int kPrepOut = 128, kPrepRowsPerCta = 1, kPrepMaxSide = 8192;

enum PrepMode : int { kPrepCopy = 0, kPrepBox2 = 1, kPrepBoxK = 2, kPrepFrac = 3 };

struct AreaTabHost {
    int mode = kPrepCopy, k = 1, taps = 1;     // taps = table row length
    float inv = 1.f;                           // 1.f / (k*k) for kPrepBoxK
    std::vector<int> start, cnt;               // [128] first source index and number of taps per output index
    std::vector<float> alpha;                  // [128][taps]
};

// computeResizeAreaTab(ssize = S, dsize = 128, scale = S/128.) in OpenCV's order
inline AreaTabHost make_area_tab(int S) {
    AreaTabHost t;
    t.start.assign(kPrepOut, 0); t.cnt.assign(kPrepOut, 1);
    if (S == kPrepOut) {
        t.mode = kPrepCopy; t.alpha.assign(kPrepOut, 1.f);
        for (int d = 0; d < kPrepOut; d++) t.start[d] = d;
        return t;
    }
    if (S % kPrepOut == 0) {
        t.k = S / kPrepOut; t.mode = t.k == 2 ? kPrepBox2 : kPrepBoxK; t.inv = 1.f / (float)(t.k * t.k);
        for (int d = 0; d < kPrepOut; d++) { t.start[d] = d * t.k; t.cnt[d] = t.k; }
        t.alpha.assign(kPrepOut, 1.f);
        return t;
    }
    t.mode = kPrepFrac;
    const double scale = (double)S / kPrepOut;
    t.taps = (int)std::ceil(scale) + 2;
    t.alpha.assign((size_t)kPrepOut * t.taps, 0.f);
    for (int dx = 0; dx < kPrepOut; dx++) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, S - fsx1);
        int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
        sx2 = std::min(sx2, S - 1);
        sx1 = std::min(sx1, sx2);
        int n = 0, first = sx1;
        float* a = &t.alpha[(size_t)dx * t.taps];
        if (sx1 - fsx1 > 1e-3) { first = sx1 - 1; a[n++] = (float)((sx1 - fsx1) / cell); }
        for (int sx = sx1; sx < sx2; sx++) a[n++] = float(1.0 / cell);
        if (fsx2 - sx2 > 1e-3) a[n++] = (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell);
        t.start[dx] = first; t.cnt[dx] = n;         // the taps are consecutive source indices first .. first+n-1
    }
    return t;
}

struct PrepParams {
    const uint8_t* frames;        // [n][fh][fw][3] BGR
    uint8_t* out;                 // [n][128][128]
    const int* start;             // device copies of the table
    const int* cnt;
    const float* alpha;
    int fh, fw, x0, y0;           // frame size and crop origin
    int mode, k, taps;
    float inv;
};

__device__ __forceinline__ int bgr2gray(const uint8_t* p) {
    return (3735 * (int)__ldg(p) + 19235 * (int)__ldg(p + 1) + 9798 * (int)__ldg(p + 2) + (1 << 14)) >> 15;
}

__global__ void __launch_bounds__(kPrepOut)
preprocess_bgr_kernel(const PrepParams P)
{
    const int dx = threadIdx.x;
    const uint8_t* frame = P.frames + (size_t)blockIdx.y * P.fh * P.fw * 3;
    uint8_t* out = P.out + (size_t)blockIdx.y * kPrepOut * kPrepOut;
    const int sx0 = P.start[dx], nx = P.cnt[dx];
    const float* ax = P.alpha + (size_t)dx * P.taps;

#pragma unroll 1
    for (int r = 0; r < kPrepRowsPerCta; r++) {
        const int dy = blockIdx.x * kPrepRowsPerCta + r;
        const int sy0 = P.start[dy], ny = P.cnt[dy];
        const uint8_t* src = frame + ((size_t)(P.y0 + sy0) * P.fw + (P.x0 + sx0)) * 3;
        int v;
        if (P.mode == kPrepFrac) {
            const float* ay = P.alpha + (size_t)dy * P.taps;
            float sum = 0.f;
            for (int j = 0; j < ny; j++) {
                const uint8_t* row = src + (size_t)j * P.fw * 3;
                float buf = 0.f;
                for (int i = 0; i < nx; i++) buf = __fadd_rn(buf, __fmul_rn((float)bgr2gray(row + 3 * i), ax[i]));
                const float term = __fmul_rn(ay[j], buf);
                sum = j == 0 ? term : __fadd_rn(sum, term);
            }
            v = min(max(__float2int_rn(sum), 0), 255);
        } else {
            int tot = 0;
            for (int j = 0; j < ny; j++) {
                const uint8_t* row = src + (size_t)j * P.fw * 3;
                for (int i = 0; i < nx; i++) tot += bgr2gray(row + 3 * i);
            }
            if (P.mode == kPrepBox2)      v = (tot + 2) >> 2;
            else if (P.mode == kPrepBoxK) v = min(max(__float2int_rn(__fmul_rn((float)tot, P.inv)), 0), 255);
            else                          v = tot;
        }
        out[dy * kPrepOut + dx] = (uint8_t)v;
    }
}

}  // namespace cnnacc
