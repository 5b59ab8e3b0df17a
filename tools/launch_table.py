"""Markdown table of a launch list written by tools/traffic_json.py (profiles/<tag>_launches_one_step.csv):
python tools/launch_table.py profiles/r2_launches_one_step.csv"""
import collections
import csv
import re
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = r["Metric Value"]


def short(n):
    m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", n)
    if "direct_copy_kernel_cuda" in n:
        return "direct_copy_kernel_cuda"
    return m.group(1) if m else n[:40]


agg = collections.OrderedDict()
for d in by.values():
    def f(k):
        try:
            return float(d.get(k, "0").replace(",", ""))
        except ValueError:
            return 0.0
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    t = f("gpu__time_duration.sum")
    a[0] += 1; a[1] += t; a[2] += f("dram__bytes_read.sum"); a[3] += f("dram__bytes_write.sum")
    a[4] += f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") * t
    a[5] += f("sm__warps_active.avg.pct_of_peak_sustained_active") * t
total = sum(a[1] for a in agg.values())
print(f"total kernel time {total / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches (ncu: serialised, cold caches - compare shares)\n")
print("| kernel | launches | time ms | share | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | warps active % |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.1f}% | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} | {(a[2] + a[3]) / a[1]:.0f} | {a[4] / a[1]:.1f} | {a[5] / a[1]:.1f} |")
