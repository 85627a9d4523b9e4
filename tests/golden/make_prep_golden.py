#!/usr/bin/env python3
"""Generate tests/golden/prep_cases.npz with OpenCV, following the reference's own pre-processing lines.  Build container only.

    python tests/golden/make_prep_golden.py

software/realtime_detect.py:582-591 is inline code in the capture loop, not a function, so it cannot be imported; the eight
lines are repeated here verbatim in behaviour (centre-crop, cv2.cvtColor(BGR2GRAY), cv2.resize(INTER_AREA)) and run with
cv2 (4.13.0 here; the reference pins no version) on the seeded frames of tests/inputs.py PREP_CASES.  Only the 128x128
outputs are stored; the frames are regenerated from their seeds at test time.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
import inputs  # noqa: E402

IMG = 128


def reference_preprocess(frame):
    h, w = frame.shape[:2]
    if w > h:
        x1 = (w - h) // 2
        crop = frame[:, x1:x1 + h]
    elif h > w:
        y1 = (h - w) // 2
        crop = frame[y1:y1 + w, :]
    else:
        crop = frame
    gray = cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY)
    return cv2.resize(gray, (IMG, IMG), interpolation=cv2.INTER_AREA)


def main():
    cv2.setNumThreads(1)
    out = {"cv2_version": np.array(cv2.__version__)}
    for c in inputs.PREP_CASES:
        frames = inputs.make_frames(c["frames"], c["n"], c["h"], c["w"])
        out[c["name"]] = np.stack([reference_preprocess(np.ascontiguousarray(f)) for f in frames])
        print(f"{c['name']:14s} {c['h']}x{c['w']} n={c['n']} mean={out[c['name']].mean():.2f}")
    np.savez_compressed(os.path.join(HERE, "prep_cases.npz"), **out)
    print("wrote prep_cases.npz", os.path.getsize(os.path.join(HERE, "prep_cases.npz")), "bytes")


if __name__ == "__main__":
    main()
