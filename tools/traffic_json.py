"""profiles/<tag>_traffic.json + <tag>_launches_one_step.csv from an `ncu --csv` launch list of tools/one_step.py (STEPS=2):
keeps the launches of the second (warm) train step - everything after the first adam_kernel up to the second one - and
averages dram__bytes_read.sum + dram__bytes_write.sum per launch for the kernels bench.py reports a roofline for.
usage: python tools/traffic_json.py gpurun_out/r1d_launches.csv r1d"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LABELS = {                      # bench.py kernel label -> substring of the demangled kernel name
    "block_bwd6": ("block_bwd6_kernel<1, 1>", "block_bwd6_kernel<true, true>", "block_bwd6_kernel"),
    "block_bwd2": ("block_bwd3_kernel<0, 1>", "block_bwd2_kernel<0, 1>"),
    "block_fwd": ("block_fwd2_kernel",),
    "skip_head": ("skip_head_kernel",),
    "gemm_nt_dx": ("gemm_nt_kernel<64, 2>",),
    "gemm_nt_dZcat": ("gemm_nt_resb_kernel<0>",),
}


def main(path, tag):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    ids = []
    for r in rows:
        if not ids or ids[-1][0] != r["ID"]:
            ids.append((r["ID"], r["Kernel Name"]))
    adam = [i for i, (_, n) in enumerate(ids) if "adam_kernel" in n]
    assert len(adam) >= 2, "need two train steps in the capture"
    keep = {ids[i][0] for i in range(adam[0] + 1, adam[1] + 1)}
    step = [r for r in rows if r["ID"] in keep]
    out_csv = os.path.join(ROOT, "profiles", f"{tag}_launches_one_step.csv")
    with open(out_csv, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(step)
    by = collections.defaultdict(lambda: [0.0, set()])
    total = 0.0
    for r in step:
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"].lower()
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
            total += v
            for label, pats in LABELS.items():
                if any(p in r["Kernel Name"] for p in pats):
                    by[label][0] += v
                    by[label][1].add(r["ID"])
    prev = {}
    try:
        prev = json.load(open(os.path.join(ROOT, "profiles", "r1c_traffic.json")))
    except Exception:
        pass
    out = {
        "source": f"profiles/{tag}_launches_one_step.csv (ncu --metrics ..dram__bytes_read/write.sum, one warm train step at cfg 2, "
                  "B=16), averaged over the kernel's launches of the step (tools/traffic_json.py)",
        "step_dram_bytes": total,
        "kernels": {k: {"dram_bytes_per_launch": v[0] / len(v[1]), "launches": len(v[1])} for k, v in by.items()},
        "generation": prev.get("generation"),
    }
    with open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["kernels"], indent=1), len(keep), "launches in the step,", total / 1e9, "GB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
