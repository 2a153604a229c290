"""Per-kernel totals of an ncu launch list (every kernel, not only this library's): python tools/ncu_all_kernels.py launches.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
tot, cnt = defaultdict(float), defaultdict(int)
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
    k = re.sub(r"<.*", "", k)
    tot[k] += us
    cnt[k] += 1
total = sum(tot.values())
print(f"{sum(cnt.values())} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised)")
print(f"{'kernel':<72s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k[:72]:<72s} {cnt[k]:>8d} {v / 1e3:>10.3f} {v / cnt[k]:>9.1f} {100 * v / total:>6.1f}%")
