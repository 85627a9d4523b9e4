#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, the default bench line, smoke(), then the ncu launch list and one --set full
# capture per kernel (tools/r2_ncu.sh).  Everything lands in gpurun_out/<tag>_*; summaries are copied to profiles/ by
# tools/collect_profiles.sh.
#   tools/gpu_evidence.sh <tag>
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.txt
tail -3 $OUT/${TAG}_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.txt; tail -2 $OUT/${TAG}_smoke.txt
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_bench.err
tail -2 $OUT/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"
bash tools/r2_ncu.sh $TAG
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/${TAG}_smoke_ncu_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke_ncu.log 2>&1; echo "smoke launch list rc=$?"
ls -la $OUT | grep ${TAG}_ | tail -30
