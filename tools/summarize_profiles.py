#!/usr/bin/env python3
"""Turns the raw ncu output brought back in gpurun_out/ into the small, committed summaries under profiles/.
Usage: python tools/summarize_profiles.py <round-tag, e.g. r1>   (runs here, no GPU: reads .csv / .ncu-rep with `ncu -i`)"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6) * v


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1) * v


def launches(tag):
    src = os.path.join(GP, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(OUT, f"{tag}_launches.csv"))
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
    per_id = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        per_id[row["ID"]]["name"] = row["Kernel Name"]
        per_id[row["ID"]][row["Metric Name"]] = (row["Metric Value"], row["Metric Unit"])
    total = 0.0
    for rec in per_id.values():
        short = re.sub(r"\(.*", "", rec["name"]).replace("void ", "").strip()[:80]
        a = agg[short]
        a["n"] += 1
        if "gpu__time_duration.sum" in rec:
            ms = to_ms(*rec["gpu__time_duration.sum"])
            a["ms"] += ms
            total += ms
        if "dram__bytes_read.sum" in rec:
            a["rd"] += to_bytes(*rec["dram__bytes_read.sum"])
        if "dram__bytes_write.sum" in rec:
            a["wr"] += to_bytes(*rec["dram__bytes_write.sum"])
    with open(os.path.join(OUT, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# {tag}: every kernel launch of ONE ResNet-50 train step (batch 256, 3x224x224, TF32), `ncu --metrics gpu__time_duration.sum,"
                "dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --replay-mode application` around `tools/ncu_step.py`\n\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes.\n\n"
                f"Total device time of the step's kernels: {total:.3f} ms in {sum(a['n'] for a in agg.values())} launches.\n\n"
                "| kernel | launches | total ms | share | avg us | DRAM read MB/launch | DRAM write MB/launch |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            f.write(f"| `{k}` | {a['n']} | {a['ms']:.3f} | {a['ms'] / total * 100:.1f}% | {a['ms'] / a['n'] * 1e3:.1f} | "
                    f"{a['rd'] / a['n'] / 1e6:.1f} | {a['wr'] / a['n'] / 1e6:.1f} |\n")
    traffic = {k: {"launches": a["n"], "ms": a["ms"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"]} for k, a in agg.items()}
    with open(os.path.join(OUT, f"{tag}_traffic.json"), "w") as f:
        json.dump({"total_ms": total, "kernels": traffic}, f, indent=1, sort_keys=True)


def full(tag, name):
    rep = os.path.join(GP, f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_ncu_full_{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, kernels matching '{name}', one ResNet-50 step (tools/ncu_step.py)\n")
        for r in rows[2:]:
            f.write("----\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:75s} {r[i][:110]} {units[i]}\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    for name in ("umma", "bn", "stem"):
        full(tag, name)
    sp = os.path.join(GP, "step_profile.tsv")
    if os.path.exists(sp):
        shutil.copy(sp, os.path.join(OUT, f"{tag}_step_profile.tsv"))
    print(sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
