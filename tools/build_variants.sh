#!/bin/bash
# Builds tuning variants of libcnnacc.so into build/variants/ (warp split of the fused kernel).
#   tools/build_variants.sh "16 4" "12 8" "16 4 1" ...      ("L0 warps, epilogue warps[, how many L0 warps use dp4a]")
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  set -- $v
  dp=${3:-8}
  out=build/variants/libcnnacc_l0w$1_epw$2_dp4a$dp.so
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function -cudart static \
       -DCNNACC_L0_WARPS=$1 -DCNNACC_EPI_WARPS=$2 -DCNNACC_L0_DP4A_WARPS=$dp -Xptxas -v -shared -o $out fpga-cnn-object-detection-accelerator_b200/csrc/cnnacc_api.cu 2>&1 \
       | grep -A2 conv_stack_fused | grep -E "registers|spill" | tr '\n' ' '
  echo " -> $out"
done
