#!/usr/bin/env python3
"""Run the -DCNNACC_TRACE build on a small device batch (CTA 0 prints its schedule trace)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import fpga_cnn_b200 as fc
wt = np.fromfile(os.path.join(ROOT, "tests", "golden", "weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0)
acc.load_weights(wt)
n = 148 * int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 6
imgs = torch.randint(0, 256, (n, 128, 128), dtype=torch.uint8, device="cuda")
if len(sys.argv) > 2 and sys.argv[2] == "infer":          # the kTail instantiation: role 4 = tail warp 0
    import inputs
    acc.load_classifier(*inputs.make_fc())
    acc.infer_batch(imgs)
else:
    acc.run_batch(imgs)
acc.synchronize()
torch.cuda.synchronize()
