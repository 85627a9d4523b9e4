#!/usr/bin/env python3
"""Merge the per-role TRACE lines of a -DCNNACC_TRACE run into one timeline (cycles since the first event)."""
import sys
ev = [l.split() for l in open(sys.argv[1]) if l.startswith("TRACE")]
names = {0: "MMA", 1: "EPI", 2: "L0a", 3: "L0b", 4: "TAIL", 5: "BACK"}
seqs = {}
for _, r, c, t in ev:
    seqs.setdefault(int(r), []).append((int(c), int(t)))
allev = []
for r, seq in seqs.items():
    base, prev = 0, None
    for c, t in seq:
        if prev is not None and t < prev - (1 << 23):
            base += 1 << 24
        prev = t
        allev.append((t + base, r, c))
allev.sort()
t0 = allev[0][0]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
for t, r, c in allev:
    if lo <= t - t0 <= hi:
        print(f"{t - t0:8d} {'              ' * r}{names.get(r, r)}:{c}")
