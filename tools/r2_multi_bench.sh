#!/bin/bash
# bench lines at N = 8 / 4 / 2 / 1 on one multi-GPU box (run under `gpurun --gpus 8`)
set -u
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29620 + n)) bench.py --gpus $n --steps 50 --warmup 3 > $OUT/r2g_bench_n$n.json 2> $OUT/r2g_bench_n$n.err; echo "bench n$n rc=$?"
done
timeout 200 python bench.py --steps 50 --warmup 3 > $OUT/r2g_bench_n1.json 2> $OUT/r2g_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    d = json.loads(open(f"gpurun_out/r2g_bench_n{n}.json").read().strip().splitlines()[-1])
    print("N=%d value %.2fM e2e %.2fM (%.2f of raw ceiling %.2fM) stream_1m %.2fM host_fed %.2fM sustained %.2fM" % (n, d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["frac_of_raw_copy_ceiling"], d["e2e"]["raw_copy_ceiling_images_per_s"]/1e6, d["stream_1m"]["images_per_s"]/1e6, d["stream_1m"]["host_fed"]["images_per_s"]/1e6, d["sustained"]["images_per_s"]/1e6), d["stream_1m"]["oracle_spot_check"])
PY
