#!/bin/bash
# usage: tools/e2e_stream_sweep.sh > gpurun_out/e2e_stream_sweep.txt
for mb in auto 8 16 32 64; do
  if [ $mb = auto ]; then timeout 120 python tools/e2e_stream_sweep.py 4096
  else CNNACC_HOST_CHUNK_MB=$mb timeout 120 python tools/e2e_stream_sweep.py 4096; fi
done
timeout 120 python tools/e2e_stream_sweep.py 1024
timeout 120 python tools/e2e_stream_sweep.py 16384
