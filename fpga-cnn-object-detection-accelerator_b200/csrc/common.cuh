// common.cuh -- small device helpers shared by the conv-stack kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cnnacc {

// Layer geometry of the reference network (arm_cnn.c:164-168): in-ch, out-ch.
constexpr int kInCh[3]  = {1, 16, 32};
constexpr int kOutCh[3] = {16, 32, 64};
constexpr int kWeightOffset[3] = {0, 144, 4752};   // byte offsets in weights.bin (arm_cnn.c:169-173)

// u8 activations x s8 weights, 4-way dot product with s32 accumulate (arm_cnn.c:106).
// __dp4a() only has same-sign overloads, so the mixed-sign form is inline PTX (SASS: IDP.4A.U8.S8).
__device__ __forceinline__ int dp4a_u8s8(uint32_t act, uint32_t wgt, int acc) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(act), "r"(wgt), "r"(acc));
    return d;
}

// arm_cnn.c:127-135 applied to two accumulators at once: v>0 ? v>>s : 0, saturate at 255, packed as
// bytes {lo, hi} in the low half-word, with `upper` moved into the high half-word.
// An arithmetic shift keeps negatives negative, and the saturating u8 convert clamps them to 0, so
// shift-then-saturate equals the reference's relu-then-shift-then-clamp.  SASS: SHF.R.S32 + I2IP.U8.S32.SAT.
__device__ __forceinline__ uint32_t act_pack2(int lo, int hi, int shift, uint32_t upper) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi >> shift), "r"(lo >> shift), "r"(upper));
    return d;
}
// Four accumulators -> one word, byte i = act(v_i).  cvt.pack gives d = (c << 16) | (sat(a) << 8) | sat(b),
// so the upper pair is packed first and rides in through c.
__device__ __forceinline__ uint32_t act_pack4(int v0, int v1, int v2, int v3, int shift) {
    return act_pack2(v0, v1, shift, act_pack2(v2, v3, shift, 0u));
}
__device__ __forceinline__ int max4(int a, int b, int c, int d) { return max(max(a, b), max(c, d)); }

// IEEE round-to-nearest a / b for operands where a is OFTEN EXACTLY ZERO (ReLU outputs, empty pooling bins).  The
// compiler's division tests its operands with FCHK and sends the whole warp through a ~45-instruction subroutine when any
// lane's dividend is zero: ncu showed every one of the tail's 20 divisions per thread taking it (profiles/
// r2_fusedtail_v1_*: 11 k of 21 k clk per image).  A harmless dividend for those lanes keeps the warp on the short
// inline path; no result changes.
__device__ __forceinline__ float div_rn_zero_ok(float a, float b) {
    const float q = __fdiv_rn(a == 0.f ? b : a, b);
    return a == 0.f ? 0.f : q;
}

}  // namespace cnnacc
