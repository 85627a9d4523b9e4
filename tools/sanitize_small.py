#!/usr/bin/env python3
"""Small workload for compute-sanitizer: 300 images through the fused kernel (multi-image CTAs), the register protocol
(dump path), infer_one (zero-copy) and the tail kernel, each checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import fpga_cnn_b200 as fc, inputs, oracle
wt = np.fromfile(os.path.join(ROOT, "tests", "golden", "weights.bin"), dtype=np.uint8)
acc = fc.CNNAccelerator(device=0)
acc.load_weights(wt); acc.set_shifts(7, 10, 11)
port = oracle.load_port()
imgs = inputs.make_images(("rng", 7), 300)
want = oracle.port_infer_batch(port, imgs, wt, (7, 10, 11))
got = acc.run_batch(imgs).reshape(300, 64, 256)
assert np.array_equal(got, want)
acc.load_image(imgs[0]); acc.start_inference(); acc.wait_done(5.0)
assert np.array_equal(acc.read_layer2_output(), want[0])
acc.read_feature_map(3, 4096); acc.read_feature_map(20, 1024)
f, _, _ = acc.infer_one(imgs[1]); assert np.array_equal(f, want[1])
fw, fb = inputs.make_fc(); acc.load_classifier(fw, fb)
cls, probs, bbox = acc.infer_batch(imgs[:64])
print("sanitize workload ok", cls[:4], bbox[0])
