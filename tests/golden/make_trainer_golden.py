#!/usr/bin/env python3
"""Generate tests/golden/trainer_case.npz from the REFERENCE's own classifier trainer.  Build container only.

    python tests/golden/make_trainer_golden.py

Runs, unmodified, `train_linear_classifier` of /root/reference/software/retrain_classifier.py:24-124 (imported; numpy only)
on the seeded training set of tests/inputs.py make_training_set and stores the (6,1024) weights / bias it returns.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import inputs  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_retrain", "/root/reference/software/retrain_classifier.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    x, y = inputs.make_training_set()
    out = {}
    for tag, kw in (("e400", dict(lr=0.01, epochs=400)), ("e1000_lr05", dict(lr=0.05, epochs=1000))):
        W, b = mod.train_linear_classifier(x, y, 6, verbose=True, **kw)
        assert W.shape == (6, 1024) and b.shape == (6,)
        out[tag + "_W"], out[tag + "_b"] = W.astype(np.float32), b.astype(np.float32)
        print(tag, "train accuracy of the returned weights:", float(((x @ W.T + b).argmax(1) == y).mean()))
    np.savez_compressed(os.path.join(HERE, "trainer_case.npz"), **out)
    print("wrote trainer_case.npz", os.path.getsize(os.path.join(HERE, "trainer_case.npz")), "bytes")


if __name__ == "__main__":
    main()
