"""Parity checkers for the conv-stack hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Two C checkers and one numpy checker:

  load_port()  -> oracle/liboracle.so, our re-entrant C restatement (cnn_oracle.c)
  load_ref()   -> oracle/_ref/arm_cnn.so, the reference's own arm_cnn.c compiled in place by
                  oracle/Makefile (None when it has not been built and cannot be: the GPU box
                  has no /root/reference, it uses the prebuilt file that travels with the repo)
  np_oracle    -> numpy restatement incl. the classifier / CAM tail

Both C libraries are called with the reference's own ctypes convention
(realtime_detect.py:389-391: argtypes=[c_void_p]*4, restype=c_int).
"""
import ctypes
import os
import subprocess

import numpy as np

from . import np_oracle  # noqa: F401

_DIR = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_DIR, "liboracle.so")
REF_LIB = os.path.join(_DIR, "_ref", "arm_cnn.so")
REF_SRC = "/root/reference/software/arm_cnn.c"


def build(quiet=True):
    """Run oracle/Makefile (compiles the restatement, and the reference when its source is present)."""
    r = subprocess.run(["make", "-C", _DIR], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    if not quiet:
        print(r.stdout)


def _needs_build(lib, src):
    return not os.path.exists(lib) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(lib))


def load_port():
    if _needs_build(PORT_LIB, os.path.join(_DIR, "cnn_oracle.c")):
        build()
    lib = ctypes.CDLL(PORT_LIB)
    lib.oracle_cnn_infer.argtypes = [ctypes.c_void_p] * 4
    lib.oracle_cnn_infer.restype = ctypes.c_int
    lib.oracle_cnn_infer_hw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.oracle_cnn_infer_hw.restype = ctypes.c_int
    lib.oracle_cnn_infer_batch.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.oracle_cnn_infer_batch.restype = ctypes.c_int
    return lib


def load_ref():
    """The reference's own compiled arm_cnn.c, or None."""
    if not os.path.exists(REF_LIB):
        if not os.path.exists(REF_SRC):
            return None
        build()
    lib = ctypes.CDLL(REF_LIB)
    lib.cnn_infer.argtypes = [ctypes.c_void_p] * 4
    lib.cnn_infer.restype = ctypes.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ref_infer(lib, image, weights_bin, shifts):
    """Call the reference cnn_infer exactly as ARMEngine.run does (realtime_detect.py:427-431)."""
    img = np.ascontiguousarray(image, dtype=np.uint8).reshape(-1)
    wt = np.ascontiguousarray(weights_bin, dtype=np.uint8)
    sh = np.array(shifts, dtype=np.int32)
    out = np.zeros(64 * 256, dtype=np.uint8)
    rc = lib.cnn_infer(_p(img), _p(wt), _p(sh), _p(out))
    if rc != 0:
        raise RuntimeError(f"reference cnn_infer returned {rc}")
    return out.reshape(64, 256)


def port_infer(lib, image, weights_bin, shifts, H=128, W=128, dump=False):
    img = np.ascontiguousarray(image, dtype=np.uint8).reshape(-1)
    assert img.size == H * W
    wt = np.ascontiguousarray(weights_bin, dtype=np.uint8)
    sh = np.array(shifts, dtype=np.int32)
    out = np.zeros(64 * (H // 8) * (W // 8), dtype=np.uint8)
    if dump:
        l0 = np.zeros(16 * (H // 2) * (W // 2), dtype=np.uint8)
        l1 = np.zeros(32 * (H // 4) * (W // 4), dtype=np.uint8)
        rc = lib.oracle_cnn_infer_hw(_p(img), H, W, _p(wt), _p(sh), _p(out), _p(l0), _p(l1))
    else:
        rc = lib.oracle_cnn_infer_hw(_p(img), H, W, _p(wt), _p(sh), _p(out), None, None)
    if rc != 0:
        raise ValueError(f"oracle_cnn_infer_hw returned {rc}")
    out = out.reshape(64, -1)
    if dump:
        return out, l0.reshape(16, H // 2, W // 2), l1.reshape(32, H // 4, W // 4)
    return out


def port_infer_batch(lib, images, weights_bin, shifts, H=128, W=128):
    imgs = np.ascontiguousarray(images, dtype=np.uint8).reshape(-1, H * W)
    n = imgs.shape[0]
    wt = np.ascontiguousarray(weights_bin, dtype=np.uint8)
    sh = np.array(shifts, dtype=np.int32)
    out = np.zeros((n, 64, (H // 8) * (W // 8)), dtype=np.uint8)
    rc = lib.oracle_cnn_infer_batch(_p(imgs), n, H, W, _p(wt), _p(sh), _p(out))
    if rc != 0:
        raise ValueError(f"oracle_cnn_infer_batch returned {rc}")
    return out
