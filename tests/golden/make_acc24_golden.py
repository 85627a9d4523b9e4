#!/usr/bin/env python3
"""Generate tests/golden/acc24_cases.npz from the REFERENCE's own bit-accurate FPGA model.  Build container only.

    python tests/golden/make_acc24_golden.py

Runs, unmodified, `fpga_conv_layer` of /root/reference/training/train_cnn.py:101-116 (imported; torch is present, the
`pycocotools` import lives inside the dataset class and is never reached): conv2d -> ((out + 2^23) mod 2^24) - 2^23 ->
floor-divide by 2^shift -> clamp 0..255 -> 2x2 max-pool, three layers, on the seeded cases of tests/inputs.py ACC24_CASES.
Tensors are float64 so every sum (|acc| <= 9.4 M) is exact.  The function clamps weights to +-127, so the cases use
weights inside that range (a case marked clamp127 clamps -128 bytes first, exactly as the trainer's exporter would have).
Asserts, before writing: the *_wrap cases differ from the int32 (arm_cnn.c) result, the *_nowrap cases equal it.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
import oracle  # noqa: E402
from oracle import np_oracle  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_train_cnn", "/root/reference/training/train_cnn.py")
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    assert tc.ACCUM_BITS == 24
    shipped = np.fromfile(os.path.join(HERE, "weights.bin"), dtype=np.uint8)
    port = oracle.load_port()
    out = {}
    for case in inputs.ACC24_CASES:
        wt = inputs.make_weights(case["weights"], shipped)
        kern = np_oracle.unpack_weights(wt)
        if case.get("clamp127"):
            kern = [np.maximum(k, -127) for k in kern]
            wt = inputs.pack_weights(kern)
        assert all(int(k.min()) >= -127 for k in kern), "fpga_conv_layer clamps weights to +-127"
        imgs = inputs.make_images(case["images"], case["n"])
        x = torch.from_numpy(imgs.astype(np.float64)).unsqueeze(1)
        for k, sh in zip(kern, case["shifts"]):
            x = tc.fpga_conv_layer(x, torch.from_numpy(k.astype(np.float64)), sh, 1.0)
        got = x.numpy()
        assert np.array_equal(got, np.round(got)) and got.min() >= 0 and got.max() <= 255
        feats = got.astype(np.uint8).reshape(case["n"], 64, 256)
        plain = oracle.port_infer_batch(port, imgs, wt, case["shifts"])         # int32 accumulator (arm_cnn.c)
        differ = float((feats != plain).mean())
        if case["wraps"]:
            assert differ > 0.02, (case["name"], differ)
        else:
            assert differ == 0.0, (case["name"], differ)
        mid = float(((feats > 0) & (feats < 255)).mean())
        print(f"{case['name']:24s} differs from int32 on {100 * differ:5.1f} % of outputs; mid-range {100 * mid:5.1f} %; mean {feats.mean():6.1f}")
        out[case["name"]] = feats
    np.savez_compressed(os.path.join(HERE, "acc24_cases.npz"), **out)
    print("wrote acc24_cases.npz", os.path.getsize(os.path.join(HERE, "acc24_cases.npz")), "bytes")


if __name__ == "__main__":
    main()
