import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile("tests/golden/weights.bin", dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(2, 4, 6)
fw, fb = inputs.make_fc(); acc.load_classifier(fw, fb)
st = torch.cuda.Stream(); acc.use_stream(st.cuda_stream)
res = {}
for B in (4096, 65536):
    nb = 2 if B == 65536 else 8
    x = [torch.randint(0, 256, (B, 128, 128), dtype=torch.uint8, device="cuda") for _ in range(nb)]
    f = [torch.empty((B, 64, 16, 16), dtype=torch.uint8, device="cuda") for _ in range(nb)]
    reps = 12 if B == 65536 else 100
    for name, fn in (("conv", lambda i: acc.run_batch(x[i % nb], out=f[i % nb])),
                     ("infer", lambda i: acc.infer_batch(x[i % nb])),
                     ("infer_upsampled", lambda i: acc.infer_batch(x[i % nb], bbox="upsampled")),
                     ("classify_feats", lambda i: acc.classify_batch(f[i % nb]))):
        for i in range(3): fn(i)
        torch.cuda.synchronize(); acc.timer_start()
        for i in range(reps): fn(i)
        ms = acc.timer_stop()
        res[f"{name}_{B}"] = reps * B / (ms / 1e3)
    del x, f; torch.cuda.empty_cache()
tops, ms = acc.probe_int8_peak(50.0)
res["int8_peak_tops"] = tops; res["int8_peak_ms"] = ms
# host-fed infer_batch
acc.use_stream(None)
h = fc.alloc_host((65536, 128, 128), np.uint8); h[:] = np.random.default_rng(1).integers(0, 256, h.shape, dtype=np.uint8)
import time
acc.infer_batch(h)
t0 = time.perf_counter()
for _ in range(3): acc.infer_batch(h)
res["infer_host_65536"] = 3 * 65536 / (time.perf_counter() - t0)
acc.infer_batch(h[:4096])
t0 = time.perf_counter()
for _ in range(20): acc.infer_batch(h[:4096])
res["infer_host_4096"] = 20 * 4096 / (time.perf_counter() - t0)
print(json.dumps(res, indent=1))
