"""Summaries of ncu CSV logs for profiles/.
  python tools/summarize_ncu.py launches <launches.csv>        per-kernel launch counts, total / average time, share
  python tools/summarize_ncu.py metrics  <forward_metrics.csv>  one line per launch with every collected metric
"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    return name


def launches(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows(path):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        tot[k] += us
        cnt[k] += 1
    mine = {k: v for k, v in tot.items() if k.startswith("afft::") or "afft" in k or "gemm_bf16" in k or "_kernel" in k and "at::" not in k}
    total = sum(mine.values())
    print(f"kernels of this library: {sum(cnt[k] for k in mine)} launches, {total / 1e3:.3f} ms total "
          f"(all captured launches: {sum(cnt.values())}, {sum(tot.values()) / 1e3:.3f} ms)")
    print(f"{'kernel':<70s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, v in sorted(mine.items(), key=lambda kv: -kv[1]):
        print(f"{k:<70s} {cnt[k]:>8d} {v / 1e3:>10.3f} {v / cnt[k]:>9.1f} {100 * v / total:>6.1f}%")
    gem = sum(v for k, v in mine.items() if "gemm_bf16" in k)
    print(f"\nGEMM (all variants) share of this library's kernel time: {100 * gem / total:.1f}%")
    others = {k: v for k, v in tot.items() if k not in mine}
    if others:
        print("\nother kernels (torch: input generation, copies):")
        for k, v in sorted(others.items(), key=lambda kv: -kv[1])[:8]:
            print(f"  {k[:90]:<90s} {cnt[k]:>6d} {v / 1e3:>10.3f} ms")


def metrics(path):
    per = OrderedDict()
    for r in rows(path):
        d = per.setdefault(int(r["ID"]), {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d[r["Metric Name"] + "#unit"] = r["Metric Unit"]
    print("idx kernel grid time_us dram_read_MB dram_write_MB tensor_active_pct l2_hit_pct dram_pct sm_pct")
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, d in per.items():
        t = d.get("gpu__time_duration.sum", 0.0)
        u = d.get("gpu__time_duration.sum#unit", "ns")
        t_us = t / 1e3 if u.startswith("n") else t
        rd, wr = d.get("dram__bytes_read.sum", 0.0) / 1e6, d.get("dram__bytes_write.sum", 0.0) / 1e6
        if d.get("dram__bytes_read.sum#unit", "byte").lower().startswith("m"):
            rd, wr = rd * 1e6, wr * 1e6
        print(f"{i:3d} {d['kernel']:<52s} {d['grid']:>12s} {t_us:8.1f} {rd:8.1f} {wr:8.1f} "
              f"{d.get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 0):6.1f} "
              f"{d.get('lts__t_sector_hit_rate.pct', 0):6.1f} {d.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', 0):6.1f} "
              f"{d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):6.1f}")
        a = agg[d["kernel"]]
        a[0] += 1
        a[1] += t_us
        a[2] += rd
        a[3] += wr
    print("\nper kernel: launches, total_us, avg_us, avg dram MB (read+write), achieved GB/s")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<52s} {a[0]:3d} {a[1]:9.1f} {a[1] / a[0]:8.1f} {(a[2] + a[3]) / a[0]:8.1f} {(a[2] + a[3]) / a[1] * 1e3:8.0f}")
    print(f"total {sum(a[1] for a in agg.values()):.1f} us over {sum(a[0] for a in agg.values())} launches")


def traffic(path):
    """JSON for profiles/r01_gemm_traffic.json (read by bench.py): DRAM bytes per GEMM launch of one forward."""
    import json
    per = OrderedDict()
    for r in rows(path):
        d = per.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d[r["Metric Name"] + "#unit"] = r["Metric Unit"]
    g = [d for d in per.values() if "gemm_bf16" in d["kernel"]]
    rd = sum(d["dram__bytes_read.sum"] for d in g)
    wr = sum(d["dram__bytes_write.sum"] for d in g)
    t = sum(d["gpu__time_duration.sum"] / (1e3 if d["gpu__time_duration.sum#unit"].startswith("n") else 1.0) for d in g)
    print(json.dumps({"what": "per-launch DRAM traffic of the GEMM launches of one forward (headline config, B=256), ncu --metrics "
                              "dram__bytes_read.sum,dram__bytes_write.sum (cold cache per launch)",
                      "launches": len(g), "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
                      "traffic_bytes_per_launch_avg": (rd + wr) / max(1, len(g)), "time_us_total": t}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics, "traffic": traffic}[sys.argv[1]](sys.argv[2])
