"""conv-only throughput at batch 65536 and 4096 for the library in CNNACC_LIB_PATH, with a bit-exactness spot check."""
import sys, os, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fpga_cnn_b200 as fc, inputs, oracle
wt = np.fromfile(os.path.join(ROOT, "tests/golden/weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(7, 10, 11)
imgs = inputs.make_images(("rng", 5), 300)
ok = bool(np.array_equal(acc.run_batch(imgs).reshape(300, 64, 256), oracle.port_infer_batch(oracle.load_port(), imgs, wt, (7, 10, 11))))
acc.set_shifts(2, 4, 6)
st = torch.cuda.Stream(); acc.use_stream(st.cuda_stream)
res = {"lib": os.path.basename(os.environ.get("CNNACC_LIB_PATH", "libcnnacc.so")), "bit_exact_300": ok}
for B, reps, nb in ((65536, 10, 2), (4096, 200, 16)):
    x = [torch.randint(0, 256, (B, 128, 128), dtype=torch.uint8, device="cuda") for _ in range(nb)]
    f = [torch.empty((B, 64, 16, 16), dtype=torch.uint8, device="cuda") for _ in range(nb)]
    for i in range(3): acc.run_batch(x[i % nb], out=f[i % nb])
    torch.cuda.synchronize(); acc.timer_start()
    for i in range(reps): acc.run_batch(x[i % nb], out=f[i % nb])
    res[f"conv_{B}"] = round(reps * B / (acc.timer_stop() / 1e3) / 1e6, 3)
    del x, f; torch.cuda.empty_cache()
print(json.dumps(res))
