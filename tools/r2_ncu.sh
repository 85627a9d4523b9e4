#!/bin/bash
# ncu evidence of round 2 (under gpurun): launch list + one --set full capture per kernel of tools/ncu_workload.py
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2}
W="python tools/ncu_workload.py"
timeout 120 $W > $OUT/${TAG}_ncu_plain.log 2>&1 || { echo "workload failed"; tail -5 $OUT/${TAG}_ncu_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_ncu_launches.csv $W > $OUT/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
cap() { timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o $OUT/${TAG}_$3_prof $W > $OUT/${TAG}_ncu_$3.log 2>&1; echo "$3 rc=$?"; }
# conv_stack_fused_kernel launches alternate <0,0> (run_batch) and <0,1> (infer_batch): the 3rd is conv only, the 4th has the tail
if [ "${2:-all}" != "tailonly" ]; then
cap conv_stack_fused 2 conv
cap conv_stack_fused 3 convtail
fi
if [ "${2:-all}" != "convonly" ]; then
cap classify_bbox 1 tail
cap cam_bbox_upsampled 1 cam
fi
ls -la $OUT/${TAG}_*prof* 2>/dev/null
