"""summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total time and launch count per kernel
name, and per-kernel mean duration; usage: launch_summary.py file.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
seq = []
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void (b200lu::)?", "", name)
    tot[name] += v
    cnt[name] += 1
    seq.append((name, v))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
allt = sum(tot.values())
print(f"{len(seq)} launches, {allt / 1e3:.2f} ms of kernel time")
for k in sorted(tot, key=tot.get, reverse=True)[:top]:
    print(f"{tot[k] / 1e3:9.3f} ms {100 * tot[k] / allt:5.1f}%  {cnt[k]:6d} x {tot[k] / cnt[k]:9.1f} us  {k[:110]}")
