#!/usr/bin/env python3
"""Tuning helper (GPU box): parity of the loaded libcnnacc build vs the oracle on 256 images, then device-resident
throughput at batch 65536.  CNNACC_LIB_PATH selects the build.  Prints one line."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import fpga_cnn_b200 as fc, inputs, oracle

wt = np.fromfile(os.path.join(ROOT, "tests", "golden", "weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0)
ok = True
port = oracle.load_port()
for wkind, shifts in (("shipped", (2, 4, 6)), (("rng", 3), (9, 12, 13))):
    w = inputs.make_weights(wkind, wt)
    acc.load_weights(w); acc.set_shifts(*shifts)
    imgs = inputs.make_images(("rng", 7), 300)
    got = acc.run_batch(imgs).reshape(300, 64, 256)
    want = oracle.port_infer_batch(port, imgs, w, shifts)
    ok &= bool(np.array_equal(got, want))
acc.load_weights(wt); acc.set_shifts(2, 4, 6)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
imgs = torch.randint(0, 256, (B, 128, 128), dtype=torch.uint8, device="cuda")
feats = torch.empty((B, 64, 16, 16), dtype=torch.uint8, device="cuda")
for _ in range(3): acc.run_batch(imgs, out=feats)
acc.synchronize()
best = 1e9
for rep in range(3):
    acc.timer_start()
    for _ in range(5): acc.run_batch(imgs, out=feats)
    best = min(best, acc.timer_stop() / 5)
print(f"{os.path.basename(os.environ.get('CNNACC_LIB_PATH', 'libcnnacc.so')):34s} parity={'OK' if ok else 'FAIL'}  {best:.3f} ms/step  {B / best / 1e3:.3f} M img/s", flush=True)
