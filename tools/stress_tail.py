"""Stress of the in-kernel tail's hand-over protocol (named barriers have no time-out): many infer_batch calls with random batch
sizes, shifts and output selections, each compared with run_batch -> classify_batch.  Run under `timeout`."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile(os.path.join(ROOT, "tests/golden/weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
pool = torch.randint(0, 256, (20000, 128, 128), dtype=torch.uint8, device="cuda")
t0, calls, images = time.time(), 0, 0
while time.time() - t0 < budget:
    n = int(rng.choice([1, 2, 3, 147, 148, 149, 295, 296, 297, int(rng.integers(1, 20000))]))
    ncls = int(rng.choice([1, 6, 6, 6, 16]))
    acc.load_classifier(*inputs.make_fc(seed=int(rng.integers(0, 1000)), n_cls=ncls))
    acc.set_shifts(*[int(v) for v in rng.integers(0, 14, 3)])
    x = pool[int(rng.integers(0, 20000 - n + 1)):][:n]
    mode = int(rng.integers(0, 3))
    f = acc.run_batch(x)
    c0, p0, b0 = acc.classify_batch(f.reshape(n, 64, 256), bbox="upsampled" if mode == 2 else "vec", logits=(mode == 1))
    c1, p1, b1 = acc.infer_batch(x, bbox="upsampled" if mode == 2 else "vec", logits=(mode == 1))
    torch.cuda.synchronize()
    assert torch.equal(c0, c1) and torch.equal(p0, p1) and torch.equal(b0, b1), (n, ncls, mode)
    calls += 1; images += n
print(f"stress ok: {calls} infer_batch calls, {images} images, {time.time() - t0:.0f} s")
