"""Bring-up helper: run the fused kernel stage by stage (CNNACC_DEBUG_LEVEL) in separate processes."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import fpga_cnn_b200 as fc, inputs
wt = np.fromfile(%r + "/tests/golden/weights.bin", dtype=np.uint8)
a = fc.CNNAccelerator(); a.load_weights(wt); a.set_shifts(7, 10, 11)
imgs = inputs.make_images(("rng", 1), int(sys.argv[1]))
out = a.run_batch(imgs)
print("ok", out.shape, int(out.sum()))
''' % (ROOT, ROOT, ROOT)
for n in (1, 3):
    for level in (1, 2, 3, 4, 5, 6, 99):
        env = dict(os.environ, CNNACC_DEBUG_LEVEL=str(level))
        r = subprocess.run([sys.executable, "-c", code, str(n)], capture_output=True, text=True, env=env, timeout=120)
        tail = (r.stdout.strip().splitlines() or [""])[-1] + " | " + (r.stderr.strip().splitlines() or [""])[-1][:200]
        print(f"n={n} level={level}: rc={r.returncode} {tail}", flush=True)
