import sys
ev=[l.split() for l in open(sys.argv[1]) if l.startswith('TRACE')]
def unwrap(seq):
    out=[]; base=0; prev=None
    for c,t in seq:
        if prev is not None and t < prev: base += 1<<24
        prev=t; out.append((c,t+base))
    return out
mm=unwrap([(int(c),int(t)) for _,r,c,t in ev if r=='0'])
cc=unwrap([(int(c),int(t)) for _,r,c,t in ev if r=='1'])
t0=mm[0][1]
print("MMA warp events (code, cycles since start, delta):")
prev=t0
for c,t in mm:
    print(f"  {c:3d} {t-t0:8d} {t-prev:7d}")
    prev=t
print("consumer warp 0 events:")
prev=t0
for c,t in cc[:int(sys.argv[2]) if len(sys.argv)>2 else 70]:
    print(f"  {c:3d} {t-t0:8d} {t-prev:7d}")
    prev=t
