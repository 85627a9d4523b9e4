#!/bin/bash
# quick GPU iteration: tail + conv parity subsets, pipeline timing, optional trace build
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-q}
timeout 600 python -m pytest tests/test_tail_gpu.py tests/test_conv_gpu.py -m gpu -x -q -k "fused_tail or tail_fixtures or logits or golden_cases or batch_sizes or full_pipeline" > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt
tail -4 $OUT/${TAG}_pytest.txt
timeout 200 python tools/pipe_timing.py > $OUT/${TAG}_pipe.json 2> $OUT/${TAG}_pipe.err; echo "pipe rc=$?"; cat $OUT/${TAG}_pipe.json; tail -3 $OUT/${TAG}_pipe.err
if [ -f build/variants/libcnnacc_trace.so ]; then
  CNNACC_LIB_PATH=$PWD/build/variants/libcnnacc_trace.so python tools/trace_run.py 8 infer > $OUT/${TAG}_trace.raw 2>&1; python tools/trace_print.py $OUT/${TAG}_trace.raw > $OUT/${TAG}_trace.txt
fi
for v in build/variants/lib_*.so; do [ -f $v ] && CNNACC_LIB_PATH=$PWD/$v python tools/pipe_timing_short.py 2>&1 | tail -1; done
