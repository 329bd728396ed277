"""Opcode histogram (weighted by executed count) + top stall lines from an ncu source-page CSV."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops = collections.Counter(); tot = 0
body = []
for r in rows[2:]:
    if len(r) <= iex: continue
    try: ex = int(r[iex]); sm = int(r[isamp])
    except: continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src.split()[0]
    base = op.split(".")[0]
    if base in ("LDS","STS","LDG","STG","LDL","STL"): base = ".".join(op.split(".")[:1]) + ("." + op.split(".")[-1] if op.split(".")[-1] in ("128","64","U8","U16") else "")
    ops[base] += ex; tot += ex
    body.append((sm, ex, src))
print("total warp-instructions executed:", tot)
for op, n in ops.most_common(40):
    print("  %-14s %12d  %5.1f%%" % (op, n, 100.0 * n / tot))
print("top stall lines (samples, executed, sass):")
for sm, ex, src in sorted(body, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("  %6d %10d  %s" % (sm, ex, src[:110]))
