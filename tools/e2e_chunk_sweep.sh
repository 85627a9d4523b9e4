#!/bin/bash
# e2e (host buffers) rate of cnnacc_run_batch against staging chunk size, slot count and pipeline style, per batch size.
# Output: gpurun_out/e2e_chunk_sweep.txt
out=gpurun_out/e2e_chunk_sweep.txt; mkdir -p gpurun_out; : > $out
for B in ${BATCHES:-4096 65536}; do
 for pipe in ${PIPES:-slots engines}; do
  for sl in ${SLOTS:-3 4 8}; do
    echo "batch $B pipe $pipe slots $sl" >> $out
    for mb in ${CHUNKS:-default 2 4 8 16}; do
      if [ $mb = default ]; then env -u CNNACC_HOST_CHUNK_MB CNNACC_HOST_PIPE=$pipe CNNACC_HOST_SLOTS=$sl python tools/e2e_sweep.py $B >> $out 2>&1
      else CNNACC_HOST_PIPE=$pipe CNNACC_HOST_SLOTS=$sl CNNACC_HOST_CHUNK_MB=$mb python tools/e2e_sweep.py $B >> $out 2>&1; fi
    done
  done
 done
done
cat $out
