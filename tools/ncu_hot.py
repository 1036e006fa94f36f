"""Top stall-sample instructions from `ncu -i X.ncu-rep --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
src = hdr.index("Source")
ie = hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((float(r[si] or 0), float(r[ie] or 0), r[src]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for v, n, s in sorted(data, key=lambda t: -t[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v:8.0f} {100*v/tot:5.1f}%  exec {n:10.0f}  {s[:110]}")
