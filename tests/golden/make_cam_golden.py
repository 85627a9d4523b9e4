#!/usr/bin/env python3
"""Generate tests/golden/cam_cases.npz from the REFERENCE ITSELF.  Run in the build container only.

    python tests/golden/make_cam_golden.py

Runs, unmodified, software/pynq_inference.py's Classifier.classify and Classifier.get_cam_bbox (imported from
/root/reference; they need numpy + Pillow only -- Pillow 12.2.0 here, the reference pins no version) on the seeded feature
sets of tests/inputs.py CAM_CASES and on the conv-stack features of tests/golden/conv_cases.npz, with the seeded (6,1024)
classifier of inputs.make_fc().  Also stores raw PIL BILINEAR 16x16 -> 128x128 resizes of seeded maps (pil__*), which pin
oracle/np_oracle.pil_resize_bilinear_u8 on its own.

Per case: <name>__cls (argmax of Classifier.classify), <name>__box (n,6,4) = the box for EVERY class index,
<name>__cam (n,128,128) u8 = round(cam_full*255) for the classified class.
"""
import importlib.util
import os
import sys

import numpy as np
from PIL import Image
import PIL

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests"))
import inputs  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_pynq_inference", os.path.join(REF, "software", "pynq_inference.py"))
    pi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pi)
    fc_w, fc_b = inputs.make_fc()
    clf = pi.Classifier.__new__(pi.Classifier)          # __init__ only np.load()s the three files
    clf.weight, clf.bias, clf.num_classes, clf.class_names = fc_w, fc_b, fc_w.shape[0], None

    sets = [(c["name"], inputs.make_features(c["features"], c["n"])) for c in inputs.CAM_CASES]
    conv = np.load(os.path.join(HERE, "conv_cases.npz"))
    for name in ("rng_shipped_mid", "smooth_shipped", "rng_random_mid"):
        sets.append(("cam_conv_" + name, conv[name][:4]))

    out = {"pil_version": np.array(PIL.__version__)}
    for name, feats in sets:
        n = feats.shape[0]
        cls = np.zeros(n, dtype=np.int32)
        box = np.zeros((n, 6, 4), dtype=np.int32)
        cam = np.zeros((n, 128, 128), dtype=np.uint8)
        for i in range(n):
            cls[i] = clf.classify(feats[i])[0]
            for k in range(6):
                cam_full, b = clf.get_cam_bbox(feats[i], k)
                box[i, k] = [int(v) for v in b]
                if k == cls[i]:
                    q = np.rint(cam_full * 255.0).astype(np.uint8)
                    assert np.array_equal(q.astype(np.float32) / 255.0, cam_full)
                    cam[i] = q
        out[name + "__cls"], out[name + "__box"], out[name + "__cam"] = cls, box, cam
        print(f"{name:28s} n={n} classes {np.bincount(cls, minlength=6)} full-frame boxes "
              f"{(box == [0, 0, 127, 127]).all(-1).sum()}/{n * 6}")

    rng = np.random.default_rng(77)
    src = rng.integers(0, 256, (6, 16, 16), dtype=np.uint8)
    src[1] >>= 3
    src[2] = 255
    src[3] = np.where(rng.random((16, 16)) < 0.1, 255, 0)
    out["pil__src"] = src
    out["pil__dst"] = np.stack([np.array(Image.fromarray(s).resize((128, 128), Image.BILINEAR)) for s in src])
    np.savez_compressed(os.path.join(HERE, "cam_cases.npz"), **out)
    print("wrote cam_cases.npz", os.path.getsize(os.path.join(HERE, "cam_cases.npz")), "bytes")


if __name__ == "__main__":
    main()
