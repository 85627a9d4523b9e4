#!/bin/bash
# Round-2 baseline evidence of the round-1 tree (runs under gpurun): sustained conv-stack clocks, --set full captures of the
# two tail kernels, H2D-only link rate.
set -u
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/r2_sustained.py <<'PY'
import sys, os, time, json, threading, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import fpga_cnn_b200 as fc
import pynvml
pynvml.nvmlInit(); nh = pynvml.nvmlDeviceGetHandleByIndex(0)
wt = np.fromfile("tests/golden/weights.bin", dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(2, 4, 6)
st = torch.cuda.Stream(); acc.use_stream(st.cuda_stream)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nbuf = max(2, (2 << 30) // (B * 32768))
imgs = [torch.randint(0, 256, (B, 128, 128), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
feats = [torch.empty((B, 64, 16, 16), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
for i in range(5): acc.run_batch(imgs[i % nbuf], out=feats[i % nbuf])
torch.cuda.synchronize()
samples, stop = [], [False]
def poll():
    while not stop[0]:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(nh, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(nh) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(nh)))
        time.sleep(0.01)
th = threading.Thread(target=poll, daemon=True); th.start()
res = []
steps = int(4.0 / (B / 17e6))
for rep in range(2):
    t0 = time.time(); acc.timer_start()
    for i in range(steps): acc.run_batch(imgs[i % nbuf], out=feats[i % nbuf])
    ms = acc.timer_stop(); t1 = time.time()
    s = [(m, p, r) for t, m, p, r in samples if t0 + 0.3 <= t <= t1]
    res.append({"batch": B, "steps": steps, "seconds": ms / 1e3, "images_per_s": steps * B / (ms / 1e3),
                "sm_mhz_median": float(np.median([x[0] for x in s])), "sm_mhz_min": min(x[0] for x in s), "power_w_max": max(x[1] for x in s),
                "reasons_or": hex(int(np.bitwise_or.reduce([x[2] for x in s]))), "samples": len(s)})
stop[0] = True
# per-100-step timing drift inside one more long run
acc.timer_start()
for i in range(50): acc.run_batch(imgs[i % nbuf], out=feats[i % nbuf])
res.append({"burst_50_steps_images_per_s": 50 * B / (acc.timer_stop() / 1e3)})
print(json.dumps(res, indent=1))
PY
python /tmp/r2_sustained.py 4096 > $OUT/r2_sustained_4096.json 2> $OUT/r2_sustained_4096.err; echo "sustained rc=$?"
python /tmp/r2_sustained.py 65536 > $OUT/r2_sustained_65536.json 2>> $OUT/r2_sustained_4096.err; echo "sustained65536 rc=$?"
cat $OUT/r2_sustained_4096.json $OUT/r2_sustained_65536.json

cat > /tmp/r2_tail.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile("tests/golden/weights.bin", dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(2, 4, 6)
fw, fb = inputs.make_fc(); acc.load_classifier(fw, fb)
x = torch.randint(0, 256, (16384, 128, 128), dtype=torch.uint8, device="cuda")
for i in range(3):
    acc.infer_batch(x); acc.infer_batch(x, bbox="upsampled")
torch.cuda.synchronize()
PY
python /tmp/r2_tail.py > $OUT/r2_tail_plain.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:classify_bbox -s 2 -c 1 -f -o $OUT/r2_base_tail_prof python /tmp/r2_tail.py > $OUT/r2_ncu_tail.log 2>&1; echo "ncu tail rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cam_bbox_upsampled -s 1 -c 1 -f -o $OUT/r2_base_cam_prof python /tmp/r2_tail.py > $OUT/r2_ncu_cam.log 2>&1; echo "ncu cam rc=$?"
}
python tools/pcie_peak.py > $OUT/r2_pcie_peak_n1.txt 2>&1; cat $OUT/r2_pcie_peak_n1.txt
nvidia-smi topo -m > $OUT/r2_topo_n1.txt 2>&1; lscpu | head -25 >> $OUT/r2_topo_n1.txt
ls -la $OUT | tail -8
