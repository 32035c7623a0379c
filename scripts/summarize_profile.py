#!/usr/bin/env python
"""Summarise the ncu artefacts a gpurun call brought back into profiles/<tag>_*.

  python scripts/summarize_profile.py <tag>

reads  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_<tag>.ncu-rep   (ncu --set full capture, read with `ncu -i ... --page raw --csv`)
writes profiles/<tag>_launches.md, profiles/<tag>_kernels.md (+ the raw launch csv, gzip'd).
"""
import collections
import csv
import gzip
import io
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("scgbm::", "")


lpath = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = list(csv.reader(l for l in open(lpath) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v * 1e3 if u in ("s", "second") else v
        tot[short(r[ki])] += v
        cnt[short(r[ki])] += 1
    T = sum(tot.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
        f.write(f"{len(rows) - 1} launches, {T:.1f} ms total.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for n, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| `{n}` | {cnt[n]} | {v:.3f} | {100 * v / T:.1f}% |\n")
    with open(lpath, "rb") as src, gzip.open(os.path.join(out_dir, f"{tag}_launches.csv.gz"), "wb") as dst:
        shutil.copyfileobj(src, dst)
    print("wrote", f"profiles/{tag}_launches.md")

rpath = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rpath):
    raw = subprocess.run(["ncu", "-i", rpath, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# ncu --set full capture `{tag}` (per launch)\n\n")
        f.write("| kernel | " + " | ".join(f"{w} [{units[i]}]" for w, i in idx) + " |\n")
        f.write("|---|" + "---:|" * len(idx) + "\n")
        for r in rows[2:]:
            f.write(f"| `{short(r[ki])[:60]}` | " + " | ".join(r[i] for _, i in idx) + " |\n")
    print("wrote", f"profiles/{tag}_kernels.md")
