// weights_pack.h -- host-side, once-per-load_weights repacking of weights.bin into kernel layouts.
//
// weights.bin is [layer][ob][ic][c16][tap9] s8 (arm_cnn.c:35-59, written by train_cnn.py:174-195, read
// by layer_fsm.v:156-182).  The reference re-parses it for every image (arm_cnn.c:186); here it is
// permuted once into the operand layouts the kernels want.  Pure permutation + sign-preserving byte copies.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace cnnacc {

struct LayerShape { int ic, oc, woff; };
static const LayerShape kLayers[3] = {{1, 16, 0}, {16, 32, 144}, {32, 64, 4752}};

// byte of k[o][i][tap] inside weights.bin
inline uint8_t weight_byte(const uint8_t* wbin, int layer, int o, int i, int tap) {
    const LayerShape& L = kLayers[layer];
    return wbin[L.woff + (((o / 16) * L.ic + i) * 16 + (o % 16)) * 9 + tap];
}

// conv_direct.cuh layout: per layer [oc][ic][8] words = lo[0..2], hi[0..2], 0, 0 with
// lo[dy] = w(dy,0) | w(dy,1)<<8 | w(dy,2)<<16 and hi[dy] = lo[dy] << 8 (same taps, one pixel to the right).
// Word order inside the 8 is {lo0, lo1, lo2, hi0, hi1, hi2, 0, 0} so two 16-byte loads fetch everything.
inline void pack_direct_weights(const uint8_t* wbin, std::vector<uint32_t>& out, size_t layer_off_words[3]) {
    out.clear();
    for (int l = 0; l < 3; l++) {
        const LayerShape& L = kLayers[l];
        layer_off_words[l] = out.size();
        for (int o = 0; o < L.oc; o++)
            for (int i = 0; i < L.ic; i++) {
                uint32_t lo[3];
                for (int dy = 0; dy < 3; dy++)
                    lo[dy] = (uint32_t)weight_byte(wbin, l, o, i, dy * 3 + 0)
                           | (uint32_t)weight_byte(wbin, l, o, i, dy * 3 + 1) << 8
                           | (uint32_t)weight_byte(wbin, l, o, i, dy * 3 + 2) << 16;
                out.push_back(lo[0]); out.push_back(lo[1]); out.push_back(lo[2]);
                out.push_back(lo[0] << 8); out.push_back(lo[1] << 8); out.push_back(lo[2] << 8);
                out.push_back(0); out.push_back(0);
            }
    }
}

}  // namespace cnnacc
