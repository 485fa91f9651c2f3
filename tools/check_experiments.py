#!/usr/bin/env python3
"""Next-round A/B of the two experiments that were written at the end of round 1 without GPU time left:
MVSIM_Z_DECIMATE=1 / 2 (ZFusedDec / ZFusedDecW: decimated inverse in the fused z pass) and MVSIM_ROT_PFWARP=<rows> (prefetch warp in
rotate_attenuate).  Runs tools/time_view.py in sub-processes (the knobs are read once per process), prints the stage
times and checks that the noise-free checksums agree.  NOT a test: run it on the GPU box,
    python tools/check_experiments.py [reps]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reps = sys.argv[1] if len(sys.argv) > 1 else "12"
variants = [("default", {}), ("z decimate (b,a)", {"MVSIM_Z_DECIMATE": "1"}), ("z decimate wide", {"MVSIM_Z_DECIMATE": "2"}), ("rotate pfwarp 8", {"MVSIM_ROT_PFWARP": "8"}),
            ("rotate pfwarp 16", {"MVSIM_ROT_PFWARP": "16"}), ("rotate pfwarp 32", {"MVSIM_ROT_PFWARP": "32"})]
sums = {}
for name, env in variants:
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_view.py"), reps], env=e, capture_output=True, text=True, timeout=600)
    line = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
    print(f"{name:18s} rc={r.returncode} {line}")
    if r.returncode != 0:
        print(r.stderr[-2000:])
    m = re.search(r"checksum ([0-9.eE+-]+)", line)
    sums[name] = float(m.group(1)) if m else None
base = sums.get("default")
for name, v in sums.items():
    if base is None or v is None:
        continue
    # the Poisson draw sees slightly different noise-free intensities (float rounding of a different transform order), so the
    # mean count agrees to ~1e-4 relative, not bit for bit
    print(f"{name:18s} checksum {v:.9f}  rel. difference to default {abs(v - base) / abs(base):.2e}")
