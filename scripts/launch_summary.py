#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
python scripts/launch_summary.py gpurun_out/launches.csv [skip_first_n]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1 + skip:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)  # -> us
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:70]:70s} {cnt[k]:8d} {v:10.1f} {v / cnt[k]:9.1f} {100 * v / T:6.1f}%")
print(f"{'TOTAL':70s} {sum(cnt.values()):8d} {T:10.1f}")
