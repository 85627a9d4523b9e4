"""Raw pinned-memory PCIe rates of this box (torch copies only): H2D alone, D2H alone, both at once, per copy size.
The ceiling for the e2e (host-buffer) numbers."""
import time, torch
N = 256 << 20
h_a = torch.empty(N, dtype=torch.uint8).pin_memory(); h_b = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device='cuda'); d_b = torch.empty(N, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, size, reps=4):
    k = N // size
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(k):
            sl = slice(i * size, (i + 1) * size)
            if h2d:
                with torch.cuda.stream(s1): d_a[sl].copy_(h_a[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_b[sl].copy_(d_b[sl], non_blocking=True)
    torch.cuda.synchronize(); return reps * N / (time.perf_counter() - t0) / 1e9
run(True, True, N, 2)
for mb in (1, 4, 16, 64, 256):
    s = mb << 20
    print(f'{mb:4d} MiB copies: H2D alone {run(True, False, s):.1f} GB/s; D2H alone {run(False, True, s):.1f} GB/s; '
          f'both at once {run(True, True, s):.1f} GB/s per direction')
