// conv_fused.cuh -- fused 128x128 conv-stack kernel (placeholder until the tcgen05 kernel lands).
#pragma once
#include "common.cuh"
#include "weights_pack.h"

namespace cnnacc {

struct FusedWeights {
    bool ready = false;
};

inline int fused_load_weights(FusedWeights&, const uint8_t*) { return 0; }
inline void fused_free(FusedWeights&) {}
inline int launch_fused(const FusedWeights&, cudaStream_t, const uint8_t*, int64_t, uint8_t*, const int*, int,
                        uint8_t*, uint8_t*) { return (int)cudaErrorNotSupported; }

}  // namespace cnnacc
