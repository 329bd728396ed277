"""Per-instruction stall listing from an ncu source-page CSV (ncu -i x.ncu-rep --page source --csv --print-source sass)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 150
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
tot = 0
out = []
agg = {}
for r in rows[2:]:
    if len(r) <= iex: continue
    try: s = int(r[isamp]); ex = int(r[iex])
    except ValueError: continue
    tot += s
    a = int(r[ia], 16) - base
    vals = [(int(r[i] or 0), h) for i, h in stall]
    for v, h in vals: agg[h] = agg.get(h, 0) + v
    top = sorted(vals, reverse=True)[:2]
    out.append((a, s, ex, r[isrc].strip(), top))
print("total samples", tot, " ".join("%s=%.1f%%" % (h[6:], 100.0 * v / tot) for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 100 > tot))
for a, s, ex, src, top in out:
    if s >= thr or any(k in src for k in ("SYNCS", "UTMALDG", "STG")):
        print("%05x %6d %9d  %-60s %s" % (a, s, ex, src[:60], " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
