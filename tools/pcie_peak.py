"""Raw pinned-memory PCIe rates of this box (torch copies only): H2D alone, D2H alone, both at once, per copy size.
The ceiling for the e2e (host-buffer) numbers.  Under torchrun every rank drives its own GPU AT THE SAME TIME (barrier
before each measurement, max over ranks), which is what the N-GPU e2e legs of bench.py compete with."""
import os, time, torch
import torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 256 << 20
h_a = torch.empty(N, dtype=torch.uint8).pin_memory(); h_b = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device='cuda'); d_b = torch.empty(N, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, size, reps=4):
    k = N // size
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(k):
            sl = slice(i * size, (i + 1) * size)
            if h2d:
                with torch.cuda.stream(s1): d_a[sl].copy_(h_a[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_b[sl].copy_(d_b[sl], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    return reps * N / dt / 1e9
run(True, True, N, 2)
if rank == 0: print(f'{world} rank(s) copying concurrently; GB/s PER GPU (slowest rank), x{world} for the box')
for mb in (1, 4, 16, 64, 256):
    s = mb << 20
    a, b, c = run(True, False, s), run(False, True, s), run(True, True, s)
    if rank == 0:
        print(f'{mb:4d} MiB copies: H2D alone {a:.1f} GB/s; D2H alone {b:.1f} GB/s; both at once {c:.1f} GB/s per direction'
              + (f'   [box total {a * world:.0f} / {b * world:.0f} / {c * world:.0f}]' if world > 1 else ''))
if world > 1: dist.destroy_process_group()
