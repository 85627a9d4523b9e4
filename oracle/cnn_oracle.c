/*
 * cnn_oracle.c -- CPU restatement of the reference conv stack.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product library
 * (fpga-cnn-object-detection-accelerator_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  oracle/Makefile also builds the reference's own software/arm_cnn.c
 * into oracle/_ref/arm_cnn.so, and tests/test_oracle.py checks this restatement byte-for-byte
 * against (a) that binary when present and (b) the committed fixtures in tests/golden/ that
 * were generated from that binary and from the reference's numpy path.
 *
 * What it restates (reference file:line, all under /root/reference/software/):
 *   oracle_unpack_kernels  <- arm_cnn.c:43-59   parse_kernels   file order [ob][ic][c16][tap9]
 *   oracle_layer           <- arm_cnn.c:68-146  run_layer       pad / conv / shift+ReLU+sat / pool
 *   oracle_cnn_infer       <- arm_cnn.c:159-198 cnn_infer       3 layers 1->16->32->64
 * Differences on purpose: no static scratch (re-entrant, so it can run under threads), generic
 * H x W (the reference hard-codes 128x128; arm_benchmark.py:76-121 is the generic numpy twin),
 * optional dumps of the two intermediate maps (feature-BRAM channels 0-15 and 16-47,
 * cnn_acc_top.v:48-54).  Arithmetic is identical: u8 activation x s8 weight -> s32 accumulate,
 * v>0 ? v>>shift : 0, saturate at 255, 2x2/2 max of the post-clamp values.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_NLAYERS 3
static const int k_ic[ORACLE_NLAYERS] = {1, 16, 32};
static const int k_oc[ORACLE_NLAYERS] = {16, 32, 64};

/* weights.bin byte offset of layer L: 0 / 144 / 4752 (arm_cnn.c:169-173,187). */
static size_t layer_weight_bytes(int L) { return (size_t)k_oc[L] * k_ic[L] * 9; }

/* arm_cnn.c:43-59 -- de-interleave [ob][ic][c16][tap] into kern[o][i][tap]; bytes are s8. */
void oracle_unpack_kernels(const uint8_t *raw, int8_t *kern, int oc, int ic)
{
    for (int o = 0; o < oc; o++)
        for (int i = 0; i < ic; i++)
            for (int t = 0; t < 9; t++) {
                size_t src = ((((size_t)(o / 16) * ic + i) * 16) + (o % 16)) * 9 + t;
                kern[((size_t)o * ic + i) * 9 + t] = (int8_t)raw[src];
            }
}

/* arm_cnn.c:127-135 -- the activation: floor shift of positives, zero otherwise, clamp 255. */
static inline uint8_t act(int32_t v, int shift)
{
    int32_t s = v > 0 ? (v >> shift) : 0;
    return (uint8_t)(s > 255 ? 255 : s);
}

/*
 * arm_cnn.c:68-146 -- one layer on a [ic][H][W] u8 map -> [oc][H/2][W/2] u8 map.
 * Zero "same" padding with a centred window (arm_cnn.c:72-86,93-112): instead of building
 * a padded copy, taps that fall outside the map are skipped (they would multiply zeros).
 * Returns 0, or -1 when scratch cannot be allocated.
 */
int oracle_layer(const uint8_t *in, int ic, int H, int W,
                 const int8_t *kern, int oc, int shift, uint8_t *out)
{
    int32_t *acc = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    if (!acc) return -1;
    const int oH = H / 2, oW = W / 2;
    for (int o = 0; o < oc; o++) {
        memset(acc, 0, sizeof(int32_t) * (size_t)H * W);
        for (int i = 0; i < ic; i++) {
            const uint8_t *plane = in + (size_t)i * H * W;
            const int8_t *k = kern + ((size_t)o * ic + i) * 9;
            for (int dy = -1; dy <= 1; dy++) {
                int r0 = dy < 0 ? 1 : 0, r1 = dy > 0 ? H - 1 : H;
                for (int dx = -1; dx <= 1; dx++) {
                    int32_t kv = k[(dy + 1) * 3 + (dx + 1)];
                    if (kv == 0) continue;
                    int c0 = dx < 0 ? 1 : 0, c1 = dx > 0 ? W - 1 : W;
                    for (int r = r0; r < r1; r++) {
                        const uint8_t *src = plane + (size_t)(r + dy) * W + dx;
                        int32_t *dst = acc + (size_t)r * W;
                        for (int c = c0; c < c1; c++) dst[c] += kv * (int32_t)src[c];
                    }
                }
            }
        }
        /* arm_cnn.c:115-143: activation on each of the four, then max. */
        uint8_t *o_plane = out + (size_t)o * oH * oW;
        for (int pr = 0; pr < oH; pr++) {
            const int32_t *a0 = acc + (size_t)(2 * pr) * W, *a1 = a0 + W;
            for (int pc = 0; pc < oW; pc++) {
                uint8_t m = act(a0[2 * pc], shift), v;
                v = act(a0[2 * pc + 1], shift); if (v > m) m = v;
                v = act(a1[2 * pc], shift);     if (v > m) m = v;
                v = act(a1[2 * pc + 1], shift); if (v > m) m = v;
                o_plane[(size_t)pr * oW + pc] = m;
            }
        }
    }
    free(acc);
    return 0;
}

/*
 * arm_cnn.c:159-198 generalised to H x W (both multiples of 8).  out = [64][H/8][W/8].
 * dump_l0 ([16][H/2][W/2]) and dump_l1 ([32][H/4][W/4]) may be NULL.
 * Returns 0; -1 allocation failure; -2 bad argument (shift outside 0..31, bad H/W).
 */
int oracle_cnn_infer_hw(const uint8_t *img, int H, int W, const uint8_t *weights_bin,
                        const int *shifts, uint8_t *out, uint8_t *dump_l0, uint8_t *dump_l1)
{
    if (!img || !weights_bin || !shifts || !out) return -2;
    if (H <= 0 || W <= 0 || (H % 8) || (W % 8)) return -2;
    for (int L = 0; L < ORACLE_NLAYERS; L++)
        if (shifts[L] < 0 || shifts[L] > 31) return -2;

    size_t n0 = (size_t)16 * (H / 2) * (W / 2), n1 = (size_t)32 * (H / 4) * (W / 4);
    uint8_t *m0 = (uint8_t *)malloc(n0), *m1 = (uint8_t *)malloc(n1);
    int8_t *kern = (int8_t *)malloc(layer_weight_bytes(2));
    int rc = (m0 && m1 && kern) ? 0 : -1;

    const uint8_t *src = img;
    uint8_t *dsts[ORACLE_NLAYERS] = {m0, m1, out};
    size_t woff = 0;
    int h = H, w = W;
    for (int L = 0; L < ORACLE_NLAYERS && rc == 0; L++) {
        oracle_unpack_kernels(weights_bin + woff, kern, k_oc[L], k_ic[L]);
        woff += layer_weight_bytes(L);
        rc = oracle_layer(src, k_ic[L], h, w, kern, k_oc[L], shifts[L], dsts[L]);
        src = dsts[L];
        h /= 2; w /= 2;
    }
    if (rc == 0 && dump_l0) memcpy(dump_l0, m0, n0);
    if (rc == 0 && dump_l1) memcpy(dump_l1, m1, n1);
    free(m0); free(m1); free(kern);
    return rc;
}

/* Same signature as the reference entry point (arm_cnn.c:159-162), 128x128 only. */
int oracle_cnn_infer(const uint8_t *img, const uint8_t *weights_bin, const int *shifts, uint8_t *out)
{
    return oracle_cnn_infer_hw(img, 128, 128, weights_bin, shifts, out, NULL, NULL);
}

/* n images back to back; imgs = [n][H][W], out = [n][64][H/8][W/8]. */
int oracle_cnn_infer_batch(const uint8_t *imgs, long n, int H, int W, const uint8_t *weights_bin,
                           const int *shifts, uint8_t *out)
{
    size_t in_sz = (size_t)H * W, out_sz = (size_t)64 * (H / 8) * (W / 8);
    for (long i = 0; i < n; i++) {
        int rc = oracle_cnn_infer_hw(imgs + i * in_sz, H, W, weights_bin, shifts, out + i * out_sz, NULL, NULL);
        if (rc) return rc;
    }
    return 0;
}
