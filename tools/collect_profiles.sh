#!/bin/bash
# Condenses the artefacts of tools/gpu_evidence.sh <tag> (in gpurun_out/) into the committed files under profiles/.
#   tools/collect_profiles.sh <tag>      e.g. r2_final
set -u
TAG=$1; OUT=gpurun_out; P=profiles
for k in conv convtail tail cam; do
  [ -f $OUT/${TAG}_${k}_prof.ncu-rep ] && python tools/ncu_summary.py $OUT/${TAG}_${k}_prof.ncu-rep 16384 > $P/${TAG}_${k}_ncu_summary.txt
done
# the file bench.py reads the conv kernel's DRAM traffic from must be named *_ncu_summary.txt and mention the kernel (it does)
for f in ncu_launches.csv smoke_ncu_launches.csv pytest.txt smoke.txt; do [ -f $OUT/${TAG}_$f ] && cp $OUT/${TAG}_$f $P/${TAG}_$f; done
[ -f $OUT/${TAG}_bench.json ] && cp $OUT/${TAG}_bench.json $P/${TAG}_bench.json
[ -f $OUT/${TAG}_bench_ref.json ] && cp $OUT/${TAG}_bench_ref.json $P/${TAG}_bench_ref.json
cuobjdump -sass fpga-cnn-object-detection-accelerator_b200/libcnnacc.so | grep -oE 'UTCIMMA|UTCBAR|LDTM|UTMALDG|UTMASTG|UBLKCP|USETMAXREG\S*|IMMA\S*|IDP\S*|SYNCS\.\S+|BAR\.(SYNC|ARV)\S*|NANOSLEEP\S*' | sort | uniq -c | sort -rn > $P/${TAG}_sass_opcodes.txt
ls -la $P | grep ${TAG}
