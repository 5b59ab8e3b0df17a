"""Turn an `ncu --csv` launch list (gpu__time_duration + dram bytes + tensor activity per launch) into a per-kernel
markdown table.  usage: python tools/summarize_ncu.py launches.csv > profiles/xxx.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r".*::", "", name)
        m = row["Metric Name"]
        v = float(row["Metric Value"].replace(",", ""))
        agg[name][m] += v
        if m == "gpu__time_duration.sum":
            cnt[name] += 1
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print(f"total kernel time {tot / 1e6:.3f} ms over {sum(cnt.values())} launches (ncu: serialised, cold caches - compare shares)\n")
    print("| kernel | launches | time ms | share | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | warps active % |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t, c = a["gpu__time_duration.sum"], cnt[n]
        rd, wr = a.get("dram__bytes_read.sum", 0.0), a.get("dram__bytes_write.sum", 0.0)
        print(f"| {n} | {c} | {t / 1e6:.3f} | {t / tot:.1%} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t:.0f} | "
              f"{a.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) / c:.1f} | "
              f"{a.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0) / c:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
