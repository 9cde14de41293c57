"""Decode a FLOWTIMES_CONV_TRACE file written by tc_conv4 (cta, event, index, clock64)."""
import collections
import sys

ev = collections.defaultdict(dict)
totals = collections.defaultdict(list)
for l in open(sys.argv[1]):
    a = l.split()
    if a[0] == "2":
        totals[int(a[1])].append(int(a[3]))
        continue
    ev[(int(a[0]), int(a[1]))][int(a[2])] = int(a[3])
for j in sorted(totals):
    t = sorted(totals[j])
    print(f"branch {j}: {len(t)} CTAs, cycles min {t[0]} median {t[len(t) // 2]} max {t[-1]}")
names = {1: "mma: image full", 2: "mma: acc empty", 3: "mma: weights full", 4: "mma: stage issued", 5: "load: buffer empty",
         6: "load: done", 7: "epi: acc full", 8: "epi: done", 9: "w: stage empty", 10: "load: issued", 11: "load: landed",
         12: "mma: acc committed", 13: "mma: img committed", 14: "mma: decoded"}
for cta in (0, 1):
    ks = [k for k in ev if k[0] == cta]
    if not ks:
        continue
    t0 = min(min(ev[k].values()) for k in ks)
    print("CTA", "first" if cta == 0 else "last")
    for k in sorted(ks):
        xs = [ev[k][n] - t0 for n in sorted(ev[k])]
        print(f"  {names.get(k[1], k[1]):22s} n={len(xs):3d}", xs[:14], "... last", xs[-1])
