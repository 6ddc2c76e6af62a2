#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
launches, total device time, share.  Per-launch times under ncu are cold-cache and
serialised, so SHARES are what is compared with bench.py's live CUDA-event numbers."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
# second argument: a number of launches to skip, or "after:<kernel>" = everything after the
# last launch of <kernel> (bench.py stages all batches with k_gen_orders before the timed loops)
skip_arg = sys.argv[2] if len(sys.argv) > 2 else "0"
all_rows = [r for r in rd if len(r) > iv and r[im] == "gpu__time_duration.sum"]
if skip_arg.startswith("after:"):
    pat = skip_arg[6:]
    last = max([i for i, r in enumerate(all_rows) if pat in r[ik]] or [-1])
    skip = last + 1
else:
    skip = int(skip_arg)
rd = iter(all_rows)
agg = OrderedDict()
n = 0
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1000.0 if r[iu] in ("nsecond", "ns") else v * (1000.0 if r[iu] in ("msecond", "ms") else 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
print(f"# {n - skip} launches after skipping {skip}; total {tot:.1f} us (cold-cache, serialised under ncu)")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {c:8d} {t:12.1f} {t / c:9.2f} {t / tot:7.3f}")
