#!/bin/bash
# throughput of every build/variants/lib_*.so plus the shipped library; optional trace of libcnnacc_trace.so
for v in build/variants/lib_*.so; do [ -f $v ] && CNNACC_LIB_PATH=$PWD/$v timeout 120 python tools/pipe_timing_short.py 2>&1 | tail -1; done
timeout 120 python tools/pipe_timing_short.py | tail -1
if [ -f build/variants/libcnnacc_trace.so ]; then
  CNNACC_LIB_PATH=$PWD/build/variants/libcnnacc_trace.so timeout 60 python tools/trace_run.py 8 infer > gpurun_out/${1:-v}_trace.raw 2>&1; python tools/trace_print.py gpurun_out/${1:-v}_trace.raw > gpurun_out/${1:-v}_trace.txt
fi
