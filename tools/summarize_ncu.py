#!/usr/bin/env python3
"""Turns ncu output into the small text summaries committed under profiles/.

  launches:  python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
  kernels :  python tools/summarize_ncu.py kernels gpurun_out/prof.ncu-rep > profiles/rNN_kernels.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("mvsim::", "").replace("(int)", "")
    for a, b in (("void ", ""), ("fft_kernel<", "fft<")):
        name = name.replace(a, b)
    return name.split("(")[0][:70]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    k, m, v, u = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= v or r[m] != "gpu__time_duration.sum":
            continue
        t = float(r[v].replace(",", ""))
        t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[u], 1e-6)
        a = agg.setdefault(short(r[k]), [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | ms / launch |\n|---|---:|---:|---:|---:|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {t:.3f} | {100 * t / tot:.1f} % | {t / n:.4f} |")
    print(f"\ntotal {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches (ncu: cold cache, serialised -- compare shares, not absolutes)")


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__waves_per_multiprocessor", "waves/SM"),
        ("smsp__inst_executed.sum", "warp instructions"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    for r in data:
        print(f"### `{short(r[ix['Kernel Name']])}`\n")
        for key, label in WANT:
            if key in ix:
                print(f"- {label}: {r[ix[key]]} {units[ix[key]]}")
        top = sorted(((float(r[ix[s]] or 0), s.replace('smsp__pcsamp_warps_issue_stalled_', '')) for s in stall), reverse=True)[:5]
        tot = sum(float(r[ix[s]] or 0) for s in stall) or 1.0
        print("- top stall reasons (pc samples): " + ", ".join(f"{n} {100 * v / tot:.0f} %" for v, n in top) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
