#!/usr/bin/env python3
"""Throughput of the tiled path (images larger than 128x128) vs the per-layer kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import fpga_cnn_b200 as fc
acc = fc.CNNAccelerator(device=0); acc.load_weights(np.fromfile(os.path.join(ROOT, "tests", "golden", "weights.bin"), dtype=np.uint8))
for H in (256, 512, 1024):
    n = {256: 2048, 512: 512, 1024: 128}[H]
    imgs = torch.randint(0, 256, (n, H, H), dtype=torch.uint8, device="cuda")
    for direct in (False, True):
        for _ in range(2): acc.run_batch(imgs, direct=direct)
        acc.timer_start()
        for _ in range(3): acc.run_batch(imgs, direct=direct)
        ms = acc.timer_stop() / 3
        print(f"{H}x{H} {'per-layer kernels' if direct else 'windows through the fused kernel'}: {n/ms*1e3:,.0f} img/s ({n/ms*1e3*H*H/16384/1e6:.2f} M 128x128-equivalents/s)")
