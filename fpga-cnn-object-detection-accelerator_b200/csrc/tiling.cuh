// tiling.cuh -- images larger than 128x128 through the fused 128x128 kernel, by overlapping tiles with halo recompute.
//
// The reference handles other sizes only in its numpy twin (arm_benchmark.py:76-121, H,W-generic); the FPGA solves the
// "map does not fit on chip" problem by spatial tiling through its accumulators (layer_fsm.v:66-75,205-213).  Here a
// tile is a 128x128 window = 16x16 outputs of the three-layer stack.  Output o (one dimension) depends on input pixels
// 8o-7 .. 8o+14 and on the zero padding each layer applies at the IMAGE border.  Inside a window the kernel pads at the
// WINDOW border instead, so the outermost output row/column of a window is only right where the window border is the
// image border.  Windows therefore advance by 14 outputs and overlap by 2:
//   tile t covers outputs [14t, 14t+14), window origin g(t) = clamp(14t-1, 0, Ho-16) outputs = 8*g pixels.
// Windows therefore overlap by 2 outputs; along one dimension (Ho outputs, window origin g in outputs = 8g pixels):
//   window 0: g = 0, owns outputs 0..14; then each window starts one output before the first output it owns and owns
//   14; the last window is pushed back to g = Ho-16 and owns everything up to Ho-1.
// Cost: (16/14)^2 = 1.31x recompute in the limit (512x512: 25 windows instead of 16).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cnnacc {

constexpr int kMaxTilesPerDim = 80;                      // 8192 / 8 / 14 = 73.2

// Window plan along one dimension, built on the host and passed by value to the kernels.
struct TilePlan {
    int n;                                               // windows
    short g[kMaxTilesPerDim];                            // window origin (outputs)
    short s[kMaxTilesPerDim], e[kMaxTilesPerDim];        // owned outputs [s, e)
};

inline TilePlan make_tile_plan(int Ho) {
    TilePlan p{};
    int start = 0;
    while (start < Ho && p.n < kMaxTilesPerDim) {
        int g = start - 1;
        if (g > Ho - 16) g = Ho - 16;
        if (g < 0) g = 0;
        const int hi = (g == Ho - 16) ? 15 : 14;         // last valid local output of this window
        int end = g + hi + 1;
        if (end > Ho) end = Ho;
        p.g[p.n] = (short)g; p.s[p.n] = (short)start; p.e[p.n] = (short)end;
        p.n++;
        start = end;
    }
    return p;
}
inline int tiles_per_dim(int Ho) { return make_tile_plan(Ho).n; }

// The windows are not materialised: the fused kernel's TMA box reads window (ty, tx) of image img at pixel origin
// (8*gx - 16, 8*gy - 1) of the big image (conv_fused.cuh, window mode; origins are even, so x stays 16-byte aligned).

// Nor are the 16x16 feature tiles: the kernel's layer-2 epilogue stores every output a window computes exactly (local
// rows / columns 1..14, plus 0 and 15 at image borders) straight into the big feature map; where neighbouring windows
// overlap they write identical bytes.  The plan's [s, e) ranges document which window 'owns' an output and are what the
// CPU tests check the plan with (cnnacc_tile_plan_host).

}  // namespace cnnacc
