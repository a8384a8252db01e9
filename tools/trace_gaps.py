"""Reads the [ptt trace] lines of PTT_SWEEP_TRACE=1 from stdin and prints, per sweep call, the gaps on the compute stream."""
import sys
calls, cur = [], []
for line in sys.stdin:
    if not line.startswith("[ptt trace]"):
        continue
    if "end" in line:
        calls.append(cur); cur = []
        continue
    p = line.split()
    cur.append((int(p[3]), float(p[5]), float(p[8])))
names = {0: "K3", 1: "retry", 2: "thermo", 3: "a2", 4: "sdv_y", 5: "sdv", 6: "gw", 7: "emit", 8: "fold"}
for k, c in enumerate(calls):
    main = sorted([x for x in c if x[0] in (2, 3, 4, 5, 6, 8)], key=lambda x: x[1])
    span = max(x[1] + x[2] for x in c) - min(x[1] for x in c)
    busy = sum(x[2] for x in main)
    k3 = sum(x[2] for x in c if x[0] == 0)
    gaps = []
    for a, b in zip(main[:-1], main[1:]):
        g = b[1] - (a[1] + a[2])
        if g > 1.0:
            gaps.append(f"{names[a[0]]}->{names[b[0]]}@{a[1] + a[2]:.0f}ms:{g:.1f}")
    print(f"call {k}: span {span:.1f} main-busy {busy:.1f} K3 {k3:.1f} gaps>1ms: {' '.join(gaps)}")
    if span - busy - k3 > 25:
        for x in sorted(c, key=lambda x: x[1]):
            print(f"      {names[x[0]]:7s} start {x[1]:8.2f} dur {x[2]:7.2f}")
