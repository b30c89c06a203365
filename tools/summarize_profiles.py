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


def table(src, dst):
    """One row per distinct kernel of a multi-kernel `ncu --set full` capture (first instance of each)."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = lambda r, k: r[hdr.index(k)] if k in hdr else ""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    def nbytes(r, k):
        if k not in hdr or not r[hdr.index(k)]:
            return 0.0
        return float(r[hdr.index(k)].replace(",", "")) * scale.get(units[hdr.index(k)], 1.0)
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    seen = OrderedDict()
    for r in data:
        name = short(col(r, "Kernel Name"))
        if name in seen:
            continue
        us = float(col(r, "gpu__time_duration.sum").replace(",", "")) * tscale.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
        rd, wr = nbytes(r, "dram__bytes_read.sum"), nbytes(r, "dram__bytes_write.sum")
        seen[name] = (us, rd, wr, col(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                      col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                      col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                      col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                      col(r, "launch__registers_per_thread"), col(r, "Grid Size"), col(r, "Block Size"))
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full, one launch of every remaining kernel: {src}\n\n")
        fh.write("DRAM GB/s = (dram read + write) / duration of that launch (cold-cache, serialised under ncu).\n\n")
        fh.write("| kernel | us | DRAM read MB | DRAM write MB | DRAM GB/s | dram % | tensor % | sm % | issue % | warps % | regs | grid | block |\n")
        fh.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|\n")
        f2 = lambda x: f"{float(x):.1f}" if x not in ("", "n/a") else ""
        for name, (us, rd, wr, dp, tp, sp, ip, wp, regs, grid, block) in seen.items():
            fh.write(f"| `{name}` | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | {f2(dp)} | {f2(tp)} | "
                     f"{f2(sp)} | {f2(ip)} | {f2(wp)} | {regs} | {grid} | {block} |\n")
    print(open(dst).read())


def table_long(src, dst):
    """`ncu --metrics ... --csv` (long format: one row per metric per launch) -> one row per distinct kernel, the
    LARGEST launch of each (by duration), with how many launches of it were measured."""
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        unit = r["Metric Unit"]
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "ms": 1e3, "s": 1e6, "msecond": 1e3, "usecond": 1.0,
              "nsecond": 1e-3, "second": 1e6}.get(unit, 1.0)
        d[r["Metric Name"]] = v
    best, count = OrderedDict(), {}
    for d in per.values():
        n = d["name"]
        count[n] = count.get(n, 0) + 1
        if n not in best or d["gpu__time_duration.sum"] > best[n]["gpu__time_duration.sum"]:
            best[n] = d
    g = lambda d, k: d.get(k, 0.0)
    with open(dst, "w") as fh:
        fh.write(f"# ncu roofline metrics, every remaining kernel of one full step: {src}\n\n")
        fh.write("`ncu --metrics <list> --clock-control none` on `bench.py --steps 1 --warmup 1` (cold-cache, serialised). "
                 "One row per kernel: its longest launch; DRAM GB/s = (read + write) / duration; L2 GB/s = lts__t_bytes / duration.\n\n")
        fh.write("| kernel | launches seen | us | DRAM read MB | DRAM write MB | DRAM GB/s | dram % | L2 GB/s | tensor % | sm % | issue % | warps % | regs | grid x block |\n")
        fh.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for n, d in sorted(best.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            us = d["gpu__time_duration.sum"]
            rd, wr = g(d, "dram__bytes_read.sum"), g(d, "dram__bytes_write.sum")
            fh.write(f"| `{n}` | {count[n]} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | "
                     f"{g(d, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g(d, 'lts__t_bytes.sum') / us / 1e3:.0f} | "
                     f"{g(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{g(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{g(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{g(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g(d, 'launch__registers_per_thread'):.0f} | "
                     f"{g(d, 'launch__grid_size'):.0f} x {g(d, 'launch__block_size'):.0f} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "table_long":
        table_long(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "table":
        table(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else 0)
    else:
        rep(sys.argv[2], sys.argv[3])
