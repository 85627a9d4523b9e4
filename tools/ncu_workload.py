"""Small fixed workload for the ncu captures of round 2: conv stack, infer_batch (tail in kernel), classify_batch on features,
upsampled-CAM boxes -- 16384 images each, a few launches of every kernel."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile(os.path.join(ROOT, "tests/golden/weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0); acc.load_weights(wt); acc.set_shifts(2, 4, 6)
acc.load_classifier(*inputs.make_fc())
x = torch.randint(0, 256, (16384, 128, 128), dtype=torch.uint8, device="cuda")
for i in range(3):
    f = acc.run_batch(x)
    acc.infer_batch(x)
    cls, _, _ = acc.classify_batch(f.reshape(16384, 64, 256))
    acc.classify_batch(f.reshape(16384, 64, 256), bbox="upsampled")
torch.cuda.synchronize()
