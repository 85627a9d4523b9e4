// conv_direct.cuh -- generic H x W per-layer kernel: conv3x3 + >>shift + ReLU/saturate + 2x2 max-pool.
//
// This is the size-generic path (any H, W multiple of 8; the 256x256 / 512x512 configs) and the
// on-GPU cross-check for the fused 128x128 kernel.  One launch per layer, CUDA cores only (dp4a).
// It computes what run_layer does (/root/reference/software/arm_cnn.c:68-146) for one 32x32 tile of
// conv outputs (16x16 pooled) and 16 output channels per CTA:
//   * zero "same" padding (arm_cnn.c:72-86): tile loads outside the map are zero-filled;
//   * u8 x s8 -> s32 accumulate over all input channels and taps (arm_cnn.c:93-112);
//   * pool the raw s32 first, then one activation (valid because the activation is monotone,
//     SURVEY.md 2.3-4), v>0 ? min(v>>s,255) : 0 (arm_cnn.c:127-135).
// Activations are planar CHW u8 in HBM ([n][c][h][w], arm_cnn.c:64-65), so the layer-0/1 maps can be
// handed back unchanged as feature-BRAM channels 0-47 (cnn_acc_top.v:48-54).
//
// Per (in-channel, row) a thread builds one activation word A = bytes x-1..x+2 and gets two horizontally
// adjacent outputs from it: out(x) = dp4a(A, {w0,w1,w2,0}), out(x+1) = dp4a(A, {0,w0,w1,w2}).
#pragma once
#include "common.cuh"

namespace cnnacc {

constexpr int kDirTile   = 32;               // conv pixels per tile side
constexpr int kDirPitchW = 10;               // tile row pitch in words: cols x0-4 .. x0+35
constexpr int kDirRows   = kDirTile + 2;     // rows y0-1 .. y0+32
constexpr int kDirIcc    = 8;                // input channels staged per pass
constexpr int kDirOcb    = 16;               // output channels per CTA
constexpr int kDirWPerOI = 8;                // packed words per (oc, ic): lo[3], hi[3], 2 pad

// Packed weights for this kernel: [oc][ic][8] words, lo[dy] = w(dy,0) | w(dy,1)<<8 | w(dy,2)<<16,
// hi[dy] = lo[dy] << 8.  Built on the host by pack_direct_weights() (weights_pack.h).
__global__ void __launch_bounds__(256)
conv3x3_pool_direct_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                           const uint32_t* __restrict__ wpk, int ic, int oc, int H, int W, int shift,
                           int tiles_x, int tiles_per_img, int acc24)
{
    __shared__ uint32_t s_tile[kDirIcc][kDirRows][kDirPitchW];
    __shared__ uint4    s_w[kDirIcc][kDirOcb][2];

    // image and tile share gridDim.x: gridDim.y's 65535 limit is below the 256 x 256 layer-0 tiles of an 8192 x 8192 image
    const int img   = blockIdx.x / tiles_per_img, tile = blockIdx.x % tiles_per_img;
    const int tileY = tile / tiles_x, tileX = tile % tiles_x;
    const int ocg   = blockIdx.z;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = tileX * kDirTile, y0 = tileY * kDirTile;

    const uint8_t* in_img = in + (size_t)img * ic * H * W;

    int acc[kDirOcb][4];
#pragma unroll
    for (int o = 0; o < kDirOcb; o++) { acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0; }

    // byte offset of A inside a tile row: column (2*tx - 1) relative to x0, +4 for the aligned left halo
    const int cb  = 2 * tx + 3;
    const int cw  = cb >> 2;
    const int csh = (cb & 3) * 8;

    for (int ic0 = 0; ic0 < ic; ic0 += kDirIcc) {
        const int icn = min(kDirIcc, ic - ic0);
        __syncthreads();
        // stage the input tile, word granular (W % 4 == 0 and x0 % 4 == 0, so words never straddle the edge)
        for (int idx = threadIdx.x; idx < icn * kDirRows * kDirPitchW; idx += 256) {
            int c = idx / (kDirRows * kDirPitchW);
            int rem = idx - c * (kDirRows * kDirPitchW);
            int r = rem / kDirPitchW, wcol = rem - r * kDirPitchW;
            int gy = y0 - 1 + r, gx = x0 - 4 + wcol * 4;
            uint32_t v = 0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = *reinterpret_cast<const uint32_t*>(in_img + ((size_t)(ic0 + c) * H + gy) * W + gx);
            s_tile[c][r][wcol] = v;
        }
        for (int idx = threadIdx.x; idx < icn * kDirOcb * 2; idx += 256) {
            int c = idx / (kDirOcb * 2);
            int rem = idx - c * (kDirOcb * 2);
            int o = rem >> 1, half = rem & 1;
            const uint4* src = reinterpret_cast<const uint4*>(
                wpk + ((size_t)(ocg * kDirOcb + o) * ic + (ic0 + c)) * kDirWPerOI);
            s_w[c][o][half] = src[half];
        }
        __syncthreads();

        for (int c = 0; c < icn; c++) {
            uint32_t A[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t* row = s_tile[c][2 * ty + r];
                A[r] = __funnelshift_r(row[cw], row[cw + 1], csh);
            }
#pragma unroll
            for (int o = 0; o < kDirOcb; o++) {
                const uint4 wa = s_w[c][o][0], wb = s_w[c][o][1];   // wa = lo0 lo1 lo2 hi0, wb = hi1 hi2 - -
                acc[o][0] = dp4a_u8s8(A[0], wa.x, dp4a_u8s8(A[1], wa.y, dp4a_u8s8(A[2], wa.z, acc[o][0])));
                acc[o][1] = dp4a_u8s8(A[0], wa.w, dp4a_u8s8(A[1], wb.x, dp4a_u8s8(A[2], wb.y, acc[o][1])));
                acc[o][2] = dp4a_u8s8(A[1], wa.x, dp4a_u8s8(A[2], wa.y, dp4a_u8s8(A[3], wa.z, acc[o][2])));
                acc[o][3] = dp4a_u8s8(A[1], wa.w, dp4a_u8s8(A[2], wb.x, dp4a_u8s8(A[3], wb.y, acc[o][3])));
            }
        }
    }

    const int oH = H >> 1, oW = W >> 1;
    const int px = (x0 >> 1) + tx, py = (y0 >> 1) + ty;
    if (px < oW && py < oH) {
        uint8_t* o_img = out + (size_t)img * oc * oH * oW;
#pragma unroll
        for (int o = 0; o < kDirOcb; o++) {
            if (acc24) {     // RTL / trainer accumulator: 24-bit two's complement wrap of every sum BEFORE the pool
#pragma unroll               // (accumulator.v:15, train_cnn.py:110-111); wrapping once at the end equals wrapping every add
                for (int q = 0; q < 4; q++) acc[o][q] = (int)((unsigned)acc[o][q] << 8) >> 8;
            }
            int m = max4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
            m = min(max(m, 0) >> shift, 255);
            o_img[((size_t)(ocg * kDirOcb + o) * oH + py) * oW + px] = (uint8_t)m;
        }
    }
}

}  // namespace cnnacc
