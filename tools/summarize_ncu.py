#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: launch list shares + key counters per kernel.
usage: summarize_ncu.py <launches.csv> <tag> <report.ncu-rep> [<report.ncu-rep> ...]"""
import collections
import csv
import subprocess
import sys

launch_csv, tag, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out = []
rows = list(csv.reader(open(launch_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg.setdefault(r[ki].split("(")[0].replace("void ", "").replace("septfa::", "").replace("<unnamed>::", ""), []).append(v)
tot = sum(sum(v) for v in agg.values())
out.append(f"## Launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over four forwards of 256 x 4 s (tools/profile_round.sh)\n")
out.append("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
out.append("| kernel | launches | avg us | total ms | share |\n|---|---:|---:|---:|---:|")
for k, v in agg.items():
    out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / 1e3:.3f} | {100 * sum(v) / tot:.1f} % |")
out.append(f"| total | {sum(len(v) for v in agg.values())} | | {tot / 1e3:.3f} | |\n")

want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("launch__occupancy_limit_shared_mem", "occ. limit smem (blocks)"), ("launch__occupancy_limit_registers", "occ. limit regs (blocks)"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instructions")]
out.append(f"## Key counters ({tag}): `ncu --set full --clock-control none --import-source on`\n")
seen = set()
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    h, units = rr[0], rr[1]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("septfa::", "").replace("unnamed>::", "")
        if name in seen:
            continue
        seen.add(name)
        out.append(f"### `{name}`\n")
        out.append("| counter | value |\n|---|---|")
        for key, label in want:
            if key in h:
                i = h.index(key)
                out.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        out.append("")
print("\n".join(out))
