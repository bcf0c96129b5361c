#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/r01_launches.csv  profiles/r01_launches_c2.txt
    python scripts/ncu_summary.py report   gpurun_out/r01_attn_global8.ncu-rep profiles/r01_attn_global8.txt

`launches`: the `--metrics gpu__time_duration.sum` CSV of one steady-state bench step -> per-kernel launch count, total
device time and SHARE of the step (cold-cache, serialised times: compare shares, not absolutes).
`report`: an `ncu --set full` capture -> duration, tensor-pipe / XU / issue utilisation, DRAM traffic, registers, and
the top stall reasons.  Runs here (no GPU needed): `ncu -i` only reads the report.
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:80]


def launches(src: str, dst: str) -> None:
    rows = []
    with open(src, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), r["Grid Size"], float(r["Metric Value"]) / 1e3))
    agg = defaultdict(lambda: [0, 0.0])
    for k, _, us in rows:
        agg[k][0] += 1
        agg[k][1] += us
    total = sum(v[1] for v in agg.values()) or 1.0
    with open(dst, "w") as out:
        out.write(f"# source: {src}  ({len(rows)} launches, {total / 1e3:.3f} ms of device time)\n")
        out.write("# ncu --metrics gpu__time_duration.sum --clock-control none : per-launch times are cold-cache and serialised;\n")
        out.write("# the SHARE column is what is comparable with bench.py's CUDA-event kernel_breakdown.\n")
        out.write(f"{'kernel':90s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"{k:90s} {n:8d} {us:10.1f} {us / n:9.2f} {us / total:7.4f}\n")
        fam = defaultdict(float)
        for k, (n, us) in agg.items():
            f = "gemm" if "gemm_kernel" in k else "attention" if "attention" in k else "row kernels"
            fam[f] += us
        out.write("\n# by family\n")
        for f, us in sorted(fam.items(), key=lambda kv: -kv[1]):
            out.write(f"{f:20s} {us:10.1f} us  share {us / total:.4f}\n")


KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
]


def report(src: str, dst: str) -> None:
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# source: {src}  (ncu --set full --clock-control none --import-source on; numbers under the profiler are\n")
        out.write("# for utilisation ratios only, never bench values)\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            out.write(f"\n## {short(d.get('Kernel Name', '?'))}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    out.write(f"{k:75s} {d[k]:>16s} {u.get(k, '')}\n")
            try:
                rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"])
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tr = rd * scale.get(u["dram__bytes_read.sum"], 1) + wr * scale.get(u["dram__bytes_write.sum"], 1)
                out.write(f"{'traffic (dram read + write) per launch':75s} {tr / 1e6:16.2f} MB\n")
            except (KeyError, ValueError):
                pass
            stalls = []
            for h in hdr:
                m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
                if m and d[h] not in ("", "n/a"):
                    stalls.append((float(d[h]), m.group(1)))
            out.write("# warp stall reasons (warps stalled per issue-active cycle), top 6\n")
            for v, n in sorted(stalls, reverse=True)[:6]:
                out.write(f"  {n:30s} {v:8.3f}\n")


def traffic(src: str, dst: str, key: str) -> None:
    """Append {key: {"bytes": dram read + write per launch, "source": ...}} to the JSON table bench.py reads for
    `roofline.traffic` (key = "<tag>[<m>x<n>x<k>]" as bench.py names the kernel)."""
    import json
    import os

    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tr = sum(float(d[k]) * scale.get(u[k], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    table[key] = {"bytes": int(tr), "source": f"profiles/{os.path.basename(src).replace('.ncu-rep', '.txt')} (ncu --set full, one launch)"}
    json.dump(table, open(dst, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "traffic":
        traffic(src, dst, sys.argv[4])
    else:
        {"launches": launches, "report": report}[mode](src, dst)
    print(open(dst).read())
