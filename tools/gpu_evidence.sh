#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, a bench line, the ncu launch list and one --set full capture
# of the conv-stack kernel.  Everything lands in gpurun_out/<tag>_*; summaries are copied to profiles/ by hand.
#   tools/gpu_evidence.sh <tag> [kernel-regex] [bench args...]
set -u
TAG=${1:-run}; shift || true
KRE=${1:-conv_stack}; shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1

python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt
tail -3 $OUT/${TAG}_pytest.txt

python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err

# ncu: same short command first without ncu (must exit 0), then the launch list, then --set full of the top kernel
NCMD="python bench.py --steps 2 --warmup 3 --batch 16384 --quick --no-cpu-baseline"
$NCMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $NCMD > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/${TAG}_prof $NCMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -15
