"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAPF / UBLKCP (TMA), HMMA (mma.sync),
REDG (global reductions), in libvit3d_sm100.so.   usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "3d_vit_ensemble_b200", "libvit3d_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "LDSM", "REDG", "SYNCS"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in keys:
            if op.startswith(k):
                per[cur][k] += 1
                if k == "UTMALDG" and ".MULTICAST" in op:
                    per[cur]["UTMALDG.MULTICAST"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of {os.path.basename(lib)} (cuobjdump -sass; sm_100a).  Columns: instructions, then counts per mnemonic.")
tot = collections.Counter()
for (name, c), dm in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dm)
    short = re.sub(r"^void ", "", short)[:80]
    cols = " ".join(f"{k}={c[k]}" for k in keys + ["UTMALDG.MULTICAST"] if c[k])
    if cols:
        print(f"{short:80s} n={c['_total']:6d}  {cols}")
    tot.update(c)
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in keys + ["UTMALDG.MULTICAST"]))
