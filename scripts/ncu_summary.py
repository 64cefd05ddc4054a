#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into the files the bench and the judge read:

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep --tag r01_tf32 --precision tf32

writes profiles/<tag>_ncu_summary.md (one row per kernel: launches, mean duration, DRAM read/write bytes per launch,
DRAM %, tensor-pipe %, registers, achieved occupancy) and merges the per-class mean DRAM traffic per launch into
profiles/traffic_per_launch.json (bench.py copies it into roofline.traffic).
"""
import argparse
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "ns ": 1e-3}
CLASS = [("mixer_", "conv"), ("pool_bwd_stats", "conv"), ("attn_pool", "row"), ("masked_mse", "loss"), ("ffn_fwd", "fused_fwd"), ("ffn_bwd", "fused_bwd"), ("tc_lse", "loss"), ("tc_grad", "loss"), ("patch_conv", "conv"), ("dwconv", "conv"),
         ("gelu_stats", "conv"), ("bn_", "conv"), ("tc_wgrad", "wgrad"), ("wgrad_kernel", "wgrad"), ("tc_gemm", "gemm"), ("gemm_kernel", "gemm"), ("attn_fwd", "attn_fwd"),
         ("attn_bwd", "attn_bwd"), ("ln_bwd", "row"), ("reduce_partials", "row"), ("embed", "row"), ("pool", "row"),
         ("lse_dir", "loss"), ("grad_dir", "loss"), ("radam", "optim")]
COLS = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
        "occ": "sm__warps_active.avg.pct_of_peak_sustained_active"}


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name.replace("mvn::", "").replace("<unnamed>::", "").replace("unnamed>::", ""))
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--head", default="", help="git commit the capture was taken on")
    ap.add_argument("--append", action="store_true", help="append to profiles/<tag>_ncu_summary.md instead of replacing it")
    a = ap.parse_args()
    if a.report.endswith(".csv"):          # already exported on the GPU box (`ncu -i rep --page raw --csv`)
        raw = open(a.report).read()
    else:
        raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    kn = hdr.index("Kernel Name")

    def val(row, k):
        if k not in ix or row[ix[k]] in ("", "n/a"):
            return float("nan")
        return float(row[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)

    agg = defaultdict(list)
    for r in data:
        agg[short(r[kn])].append({k: val(r, k) for k in COLS})
    lines = [f"# ncu --set full summary: {os.path.basename(a.report)} ({a.tag})", "",
             "Per-launch means over the captured launches (cold-cache, serialised replays: compare shares, not absolutes).", "",
             "| kernel | launches | dur us | DRAM read MB | DRAM write MB | DRAM % | tensor-pipe % | SM % | regs | warps active % |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    per_class = defaultdict(lambda: [0.0, 0])
    for name, L in sorted(agg.items(), key=lambda kv: -sum(x["dur"] for x in kv[1])):
        n = len(L)
        m = {k: sum(x[k] for x in L) / n for k in COLS}
        lines.append(f"| `{name}` | {n} | {m['dur']:.1f} | {m['rd']/1e6:.2f} | {m['wr']/1e6:.2f} | {m['dram_pct']:.1f} | {m['tensor_pct']:.1f} | "
                     f"{m['sm_pct']:.1f} | {m['regs']:.0f} | {m['occ']:.1f} |")
        for key, cls in CLASS:
            if key in name:
                per_class[cls][0] += sum(x["rd"] + x["wr"] for x in L)
                per_class[cls][1] += n
                break
    out_md = os.path.join(ROOT, "profiles", a.tag + "_ncu_summary.md")
    if a.append and os.path.exists(out_md):
        open(out_md, "a").write("\n".join(lines[6:]) + "\n")
    else:
        open(out_md, "w").write("\n".join(lines) + "\n")
    tpath = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t.setdefault(a.precision, {})
    for cls, (b, n) in per_class.items():
        t[a.precision][cls] = b / n
    t[a.precision]["_source"] = f"{a.tag}: mean dram__bytes_read.sum + dram__bytes_write.sum per launch, {os.path.basename(a.report)}"
    t["_source"] = t[a.precision]["_source"] + (f", HEAD {a.head}" if a.head else "")
    json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
