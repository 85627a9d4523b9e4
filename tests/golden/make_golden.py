#!/usr/bin/env python3
"""Generate tests/golden/* from the REFERENCE ITSELF.  Run in the build container only.

    python tests/golden/make_golden.py

Needs /root/reference (read-only).  What it runs, unmodified:
  * software/arm_cnn.c compiled in place by oracle/Makefile -> oracle/_ref/arm_cnn.so (cnn_infer)
  * software/dump_arm_features.py   parse_kernels, numpy_infer      (imported)
  * software/arm_benchmark.py       parse_weights, arm_conv_layer   (imported; H,W-generic => 256x256 case)
  * software/realtime_detect.py     classify_vec, bbox_vec          (imported)
It asserts that the C and numpy references agree byte-for-byte on every 128x128 case before
writing anything.  Inputs are regenerated from seeds by tests/inputs.py at test time, so the
fixtures hold outputs (+ the seeds/shifts that define each case) and stay small.

Files written:
  weights.bin        the shipped benchmark weight set (data file, weights/weights.bin, 23184 B)
  conv_cases.npz     features per case (+ intermediate maps for two cases)
  tail_cases.npz     classify_vec / bbox_vec outputs on the features above with a seeded (6,1024) fc
  kats.json          scalar known-answer tests restated from sim/module/*_tb.v
"""
import hashlib
import importlib.util
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402  (tests/inputs.py: seeded generators shared with the tests)
import oracle  # noqa: E402


def _imp(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, "software", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    oracle.build()
    ref = oracle.load_ref()
    assert ref is not None, "reference arm_cnn.c did not build"
    daf = _imp("dump_arm_features")
    ab = _imp("arm_benchmark")
    rd = _imp("realtime_detect")

    shutil.copyfile(os.path.join(REF, "weights", "weights.bin"), os.path.join(HERE, "weights.bin"))
    os.chmod(os.path.join(HERE, "weights.bin"), 0o644)
    shipped = np.fromfile(os.path.join(HERE, "weights.bin"), dtype=np.uint8)
    assert shipped.size == 23184

    conv = {}
    meta = []
    feats_for_tail = []
    for case in inputs.CONV_CASES:
        name = case["name"]
        wt = inputs.make_weights(case["weights"], shipped)
        imgs = inputs.make_images(case["images"], case["n"])
        sh = case["shifts"]
        # numpy reference hard-codes (2,4,6); patch its module constants per case (no source edit).
        daf.SH0, daf.SH1, daf.SH2 = sh
        kern = daf.parse_kernels(wt)
        kern_ab = ab.parse_weights(wt)
        for a, b in zip(kern, kern_ab):
            assert np.array_equal(a, b)
        out = np.zeros((case["n"], 64, 256), dtype=np.uint8)
        for i in range(case["n"]):
            c = oracle.ref_infer(ref, imgs[i], wt, sh)
            n = daf.numpy_infer(imgs[i].reshape(-1), kern)
            assert np.array_equal(c, n), f"{name}[{i}]: arm_cnn.c != numpy_infer"
            out[i] = c
        conv[name] = out
        if case.get("dump"):
            x = imgs[0].reshape(1, 128, 128)
            l0 = ab.arm_conv_layer(x, kern_ab[0], sh[0])
            l1 = ab.arm_conv_layer(l0, kern_ab[1], sh[1])
            l2 = ab.arm_conv_layer(l1, kern_ab[2], sh[2])
            assert np.array_equal(l2.reshape(64, 256), out[0])
            conv[name + "__l0"] = l0
            conv[name + "__l1"] = l1
        meta.append({"name": name, "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                     "mean": float(out.mean()), "nonzero": float((out != 0).mean()),
                     "midrange": float(((out > 0) & (out < 255)).mean())})
        if case.get("tail"):
            feats_for_tail.append((name, out))
        print(f"{name:28s} n={case['n']:3d} shifts={sh} mean={out.mean():7.2f} "
              f"mid={meta[-1]['midrange']*100:5.1f}% sha={meta[-1]['sha256'][:16]}")

    # generic-size case: arm_benchmark.arm_conv_layer is H,W-agnostic (arm_cnn.c is not).
    for case in inputs.HW_CASES:
        wt = inputs.make_weights(case["weights"], shipped)
        kern_ab = ab.parse_weights(wt)
        H, W = case["H"], case["W"]
        x = inputs.make_images(case["images"], 1, H, W)[0].reshape(1, H, W)
        for k, s in zip(kern_ab, case["shifts"]):
            x = ab.arm_conv_layer(x, k, s)
        conv[case["name"]] = x.reshape(64, -1)
        print(f"{case['name']:28s} {H}x{W} -> {x.shape} mean={x.mean():.2f}")
    np.savez_compressed(os.path.join(HERE, "conv_cases.npz"), **conv)

    # classifier / CAM tail with a seeded (6,1024) fc (the shipped fc_weight.npy is (6,64): SURVEY 2.4)
    fc_w, fc_b = inputs.make_fc()
    tail = {}
    names = rd.NAMES
    for name, feats in feats_for_tail:
        n = feats.shape[0]
        cls = np.zeros(n, dtype=np.int32)
        probs = np.zeros((n, 6), dtype=np.float32)
        bbox = np.zeros((n, 4), dtype=np.int32)
        for i in range(n):
            ci, _, _, p = rd.classify_vec(feats[i], fc_w, fc_b, names)
            cls[i] = ci
            probs[i] = p
            bbox[i] = rd.bbox_vec(feats[i], ci, fc_w)
        tail[name + "__cls"] = cls
        tail[name + "__probs"] = probs
        tail[name + "__bbox"] = bbox
        print(f"tail {name}: classes {np.bincount(cls, minlength=6)} "
              f"full-frame boxes {(bbox == [0, 0, 127, 127]).all(1).sum()}/{n}")
    np.savez_compressed(os.path.join(HERE, "tail_cases.npz"), **tail)

    kats = {
        "_source": "restated from /root/reference/sim/module/*_tb.v and sim/top/tb.v",
        "relu": [[100, 100], [-50, 0], [0, 0], [255, 255], [256, 255], [1000, 255], [-1, 0]],   # relu_tb.v:18-51
        "accumulator": {"overwrite": 500, "add": 300, "expect": 800},                           # accumulator_tb.v:28-61
        "conv_core": {"window": [10, 20, 30, 40, 50, 60, 70, 80, 90], "kernel": [1] * 9, "expect": 450},  # conv_core_tb.v:45-66
        "tb_image": "pixel[i] = (i*13+5) % 256; weights all 0 except byte 4 = 1; shifts 0",     # tb.v:501-527
        "cases": meta,
    }
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
