"""Key metrics of one `ncu --set full` capture: ncu -i rep --page raw --csv | python tools/ncu_extract.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_dim_x", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"  {w:75s} {vals[i][:60]:>20s} {units[i]}")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            v = float(vals[i].replace(",", ""))
        except ValueError:
            continue
        if v >= 0.3:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:8.2f} warps per issue")
