"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: share of total kernel time per kernel.
Usage: python tools/ncu_launch_summary.py launches.csv [topN]"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("# per-launch times under ncu are cold-cache and serialised: use the SHARE, not the absolute")
print(f"# total kernel time {total:.2f} ms over {sum(cnt.values())} launches")
print("  share  total_ms  count   avg_us  kernel")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100 * t / total:6.2f}%  {t:8.3f}  {cnt[name]:5d}  {1e3 * t / cnt[name]:7.1f}  {name[:100]}")
