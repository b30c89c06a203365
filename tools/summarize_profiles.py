#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_full_r1b.csv profiles/r1_launches_full.md [skip]
    python tools/summarize_profiles.py rep gpurun_out/prof_attn.ncu-rep profiles/r1_ncu_attn.md

`launches`: per-kernel-name totals of `gpu__time_duration.sum` (cold-cache, serialised: compare SHARES).  `skip` drops
the first launches (weight packing / warm-up).  `rep`: the roofline-relevant raw metrics of every profiled launch.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("mv::", "")


def launches(src, dst, skip=0):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    rows = rows[int(skip):]
    agg = OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as fh:
        fh.write(f"# ncu launch list summary: {src}\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised; compare shares).\n")
        fh.write(f"{len(rows)} launches after skipping {skip}; total {tot / 1e3:.3f} ms.\n\n")
        fh.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {n} | {us:.1f} | {us / tot:.3f} | {us / n:.2f} |\n")
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full summary: {src}\n\n")
        for r in data:
            name = short(r[hdr.index("Kernel Name")])
            fh.write(f"## `{name}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write(f"| {k} | {r[i]} | {units[i]} |\n")
            fh.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else 0)
    else:
        rep(sys.argv[2], sys.argv[3])
