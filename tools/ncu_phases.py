#!/usr/bin/env python3
"""Per-phase split of a barrier-phased kernel from an ncu report captured with --import-source on.

    ncu -i report.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_phases.py src.csv [warps]          # warps = CTAs x warps per CTA (default: the first instruction's count)

The SASS listing is cut at every BAR.SYNC; for each phase it prints the executed warp instructions per warp, the share of the pc
samples and the dominant opcodes, then the stall reasons of the samples.  (A warp that waits at a barrier is sampled on the first
instruction AFTER it: `barrier` samples of phase k are the wait for the slowest warp of phase k-1.)
"""
import collections
import csv
import re
import sys


def main(path, warps=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stalls = [h for h in hdr if h.startswith("stall_") and not h.endswith("(Not Issued)")]
    ph = 0
    tot, samples = collections.Counter(), collections.Counter()
    ops, why = collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
    for r in data:
        src = r[ix["Source"]].strip()
        n = int(r[ix["Instructions Executed"]] or 0)
        op = re.split(r"[ .]", re.sub(r"^@!?U?P\d+\s+", "", src))[0]
        tot[ph] += n
        ops[ph][op] += n
        samples[ph] += int(r[ix["# Samples"]] or 0)
        for h in stalls:
            why[ph][h] += int(r[ix[h]] or 0)
        if "BAR.SYNC" in src:
            ph += 1
    w = float(warps) if warps else float(int(data[0][ix["Instructions Executed"]] or 1))
    t_all, s_all = sum(tot.values()), sum(samples.values()) or 1
    for p in sorted(tot):
        print(f"phase {p}: {tot[p] / w:8.1f} instr/warp ({100 * tot[p] / t_all:4.1f} %)  samples {100 * samples[p] / s_all:4.1f} %  top: "
              + ", ".join(f"{o} {c / w:.0f}" for o, c in ops[p].most_common(12)))
    print(f"total per warp {t_all / w:.1f}")
    for p in sorted(why):
        t = sum(why[p].values()) or 1
        print(f"phase {p} stalls: " + ", ".join(f"{h[6:]} {100 * c / t:.0f} %" for h, c in why[p].most_common(7)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
