"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` output.
usage: python tools/ncu_hot_sass.py source.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {c: i for i, c in enumerate(hdr)}
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
print("total samples", tot)
for c in stalls:
    s = sum(int(r[ix[c]] or 0) for r in body)
    if s * 50 > tot:
        print(f"  {c:24s} {100.0 * s / tot:5.1f}%")
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = body[i]
    s = int(r[ix["# Samples"]] or 0)
    why = max(stalls, key=lambda c: int(r[ix[c]] or 0))
    print(f"{i:5d} {100.0 * s / tot:5.1f}%  {why:20s} {r[ix['Source']].strip()[:90]}")
