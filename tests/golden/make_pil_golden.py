#!/usr/bin/env python3
"""Generate tests/golden/pil_cases.npz from the REFERENCE's own image loader.  Build container only.

    python tests/golden/make_pil_golden.py

Runs, unmodified, `load_image_any` of /root/reference/software/pynq_inference.py:414-425 (imported; without `pynq` the module
only prints its simulation-mode notice) on lossless PNG files written from the seeded arrays of tests/inputs.py PIL_CASES, plus
its `.bin` branch.  Records the Pillow version that did the arithmetic.  Asserts that the numpy restatement
(oracle/np_oracle.py load_image_array) reproduces every output before writing.
"""
import importlib.util
import io
import contextlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
from oracle import np_oracle  # noqa: E402


def main():
    import PIL
    from PIL import Image
    spec = importlib.util.spec_from_file_location("ref_pynq_inference", "/root/reference/software/pynq_inference.py")
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    out = {"pillow_version": np.array(PIL.__version__)}
    with tempfile.TemporaryDirectory() as tmp:
        for case in inputs.PIL_CASES:
            arr = inputs.make_pil_image(case)
            path = os.path.join(tmp, case["name"] + ".png")
            Image.fromarray(arr, case["mode"]).save(path)
            assert np.array_equal(np.asarray(Image.open(path)), arr)          # PNG is lossless: the loader sees these bytes
            got = mod.load_image_any(path)
            assert got.shape == (16384,) and got.dtype == np.uint8
            assert np.array_equal(got, np_oracle.load_image_array(arr)), case["name"]
            out[case["name"]] = got
            print(f"{case['name']:20s} {case['mode']:4s} {case['h']}x{case['w']}  mean {got.mean():6.1f}")
        binp = os.path.join(tmp, "img.bin")
        inputs.tb_image().tofile(binp)
        assert np.array_equal(mod.load_image_any(binp), inputs.tb_image().reshape(-1))
    np.savez_compressed(os.path.join(HERE, "pil_cases.npz"), **out)
    print("wrote pil_cases.npz", os.path.getsize(os.path.join(HERE, "pil_cases.npz")), "bytes; Pillow", PIL.__version__)


if __name__ == "__main__":
    main()
