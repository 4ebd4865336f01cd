#!/usr/bin/env python
"""Condenses ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_TAG.csv profiles/launches_TAG.md [kernels_per_frame]
    python scripts/ncu_summary.py full     gpurun_out/frame_TAG.ncu-rep profiles/frame_TAG.md
    python scripts/ncu_summary.py gather-json profiles/STRICT.csv profiles/FAST.csv profiles/gather_ncu.json TAG
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def short(name: str) -> str:
    name = name.replace("void ", "").replace("unnamed>::", "").replace("lgcu::", "")
    return name.split("(")[0]


def launches(src, dst, per_frame=None):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v, g, b = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache and serialised: compare SHARES)", "",
           f"source: `{src}`, {len(rows)} launches", ""]
    agg = OrderedDict()
    for r in rows:
        key = short(r[k])
        t = float(r[v].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    out += ["| kernel | launches | total µs | share |", "|---|---:|---:|---:|"]
    for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {key} | {n} | {t / 1e3:.1f} | {100 * t / total:.1f}% |")
    out += ["", f"total {total / 1e3:.1f} µs over {len(rows)} launches", ""]
    if per_frame:
        out += [f"## first frame ({per_frame} launches)", "", "| # | kernel | grid | block | µs |", "|---:|---|---|---|---:|"]
        for i, r in enumerate(rows[:per_frame]):
            out.append(f"| {i} | {short(r[k])} | {r[g]} | {r[b]} | {float(r[v].replace(',', '')) / 1e3:.2f} |")
    open(dst, "w").write("\n".join(out) + "\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, s) for m, s in FULL_METRICS if m in idx]
    out = [f"# ncu --set full summary of `{src}` ({len(data)} launches; values per launch, under the profiler: for shares/ratios, not bench numbers)", "",
           "| kernel | " + " | ".join(f"{s} [{units[idx[m]]}]" if units[idx[m]] else s for m, s in cols) + " |", "|---|" + "---:|" * len(cols)]
    for r in data:
        vals = []
        for m, _ in cols:
            x = r[idx[m]]
            try:
                f = float(x.replace(",", ""))
                vals.append(f"{f:.4g}" if abs(f) < 1e6 else f"{f:.3e}")
            except ValueError:
                vals.append(x)
        out.append(f"| {short(r[idx['Kernel Name']])} | " + " | ".join(vals) + " |")
    open(dst, "w").write("\n".join(out) + "\n")




def gather_json(strict_csv, fast_csv, dst, tag):
    """ncu --metrics CSV logs of the shader-order gather (algorithmic FP32 basis) and of the throughput gather + side-pyramid pack of the
    same tree -> the small JSON bench.py reads for `roofline.traffic` and the FP32 figures (profiles/gather_ncu.json)."""
    import json

    def load(path):
        rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
        hdr = rows[0]
        k, m, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
        out = {}
        for r in rows[1:]:
            out.setdefault(short(r[k]), {})[r[m]] = float(r[v].replace(",", ""))
        return out

    def pick(d):
        return {"warp_inst": d["smsp__inst_executed.sum"], "thread_inst": d["smsp__thread_inst_executed.sum"],
                "fp32_thread_inst": d["smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"], "fadd": d["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"],
                "fmul": d["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"], "ffma": d["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"],
                "dram_bytes": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"], "time_us_under_ncu": d["gpu__time_duration.sum"] / 1e3}

    strict, fast = load(strict_csv), load(fast_csv)
    out = {"tag": tag, "workload": "3840x2160 synthetic frame (seed 0xC0FFEE), one launch each, ncu --metrics ... --clock-control none",
           "sources": [strict_csv, fast_csv]}
    for name, d in list(strict.items()) + list(fast.items()):
        key = "strict" if "Strict" in name else ("pack" if "pack" in name else "fast")
        out[key] = dict(pick(d), kernel=name)
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
    elif sys.argv[1] == "gather-json":
        gather_json(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
    else:
        full(sys.argv[2], sys.argv[3])
