"""What would overlapping consecutive host-buffer calls give?  Two handles, two host threads, each looping over synchronous
run_batch calls on its own pinned buffers: the aggregate rate is what a pipelined (asynchronous) call sequence could reach."""
import os, sys, time, threading, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fpga_cnn_b200 as fc
wt = np.fromfile(os.path.join(ROOT, "tests/golden/weights.bin"), dtype=np.uint8)
B, steps = 4096, 40
def make():
    a = fc.CNNAccelerator(device=0); a.load_weights(wt)
    x = fc.alloc_host((B, 128, 128), np.uint8); x[:] = np.random.default_rng(1).integers(0, 256, x.shape, dtype=np.uint8)
    y = fc.alloc_host((B, 64, 16, 16), np.uint8)
    a.run_batch(x, out=y)
    return a, x, y
for nthreads in (1, 2, 3):
    objs = [make() for _ in range(nthreads)]
    def work(o):
        a, x, y = o
        for _ in range(steps): a.run_batch(x, out=y)
    th = [threading.Thread(target=work, args=(o,)) for o in objs]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{nthreads} thread(s): {nthreads * steps * B / dt / 1e6:.2f} M img/s aggregate ({nthreads * steps * B * 16384 / dt / 1e9:.1f} GB/s per direction)")
    for a, _, _ in objs: a.close()
