"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/agg_launches.py file.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    out = []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            u = d["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            out.append((int(d["ID"]), re.sub(r"\(.*", "", d["Kernel Name"])[:70], d.get("Grid Size"), v))
    return out


if __name__ == "__main__":
    L = load(sys.argv[1])
    if len(sys.argv) > 3:
        L = [r for r in L if int(sys.argv[2]) <= r[0] <= int(sys.argv[3])]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, n, _, v in L:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.1f} us {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")
    print(f"{tot:9.1f} us {len(L):5d} launches")
