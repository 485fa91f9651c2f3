#!/usr/bin/env python3
"""Times the 'next' rows of SURVEY section 8f at the reference's own sizes (host-buffer entry points, wall clock incl. copies)
beside the CPU oracle: sphere phantom, bead phantom, post-acquisition chain, and the whole main() loop.
Usage (GPU box): python tools/time_next_rows.py [--no-oracle]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import mvsim_b200 as mv  # noqa: E402

S, B = mv.SimulateMultiViewDataset, mv.SimulateBeads
with_oracle = "--no-oracle" not in sys.argv
if with_oracle:
    from oracle import oracle as orc


def best(fn, reps=3):
    t = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t.append(time.perf_counter() - t0)
    return min(t), out


rows = []
ctx = mv.Context(0)
S.simulate(size=121, ctx=ctx)                                   # warm up (context, pools)

t, gt = best(lambda: S.simulate(ctx=ctx))
to = best(lambda: orc.simulate_phantom(289), 1)[0] if with_oracle else None
rows.append(("simulate() sphere phantom 289^3 (580^3 render + downSample2x)", t, to))

interval = B.interval((1024, 1024, 1024))
pts = B.transformPoints(B.randomPoints(2000, interval), [45], 0, interval)[0]
t, beads = best(lambda: B.renderPoints([pts], interval, (1.0, 1.0, 3.0), ctx=ctx)[0], 2)
to = best(lambda: orc.render_beads(pts, (1.0, 1.0, 3.0), interval[0], interval[1]), 1)[0] if with_oracle else None
rows.append(("renderPoints 2000 beads into 1023^3 (4.3 GB image incl. download)", t, to))
del beads

acq = np.ascontiguousarray(gt[::3])
t, iso = best(lambda: S.makeIsotropic(acq, 3, ctx=ctx))
to = best(lambda: orc.make_isotropic(acq, 3), 1)[0] if with_oracle else None
rows.append(("makeIsotropic 289x289x97 -> 289^3", t, to))

t, _ = best(lambda: S.rotateAroundAxis(iso, 0, -52, ctx=ctx))
to = best(lambda: orc.rotate(iso, 0, -52), 1)[0] if with_oracle else None
rows.append(("rotate back 289^3", t, to))

ws = [S.computeWeightImage(gt.shape, ctx=ctx) for _ in range(7)]
t, _ = best(lambda: S.normalizeWeights([w.copy() for w in ws], 3.0, ctx=ctx))
to = best(lambda: orc.normalize_weights([w.copy() for w in ws], 3.0), 1)[0] if with_oracle else None
rows.append(("weight normalisation, 7 views of 289^3", t, to))

t0 = time.perf_counter()
r = mv.run_main(None, ctx=ctx, log=lambda *_: None, keep=())
rows.append((f"main(): {len(r['angles'])} views of 289^3, PSF 51^3, inc 3, SNR 25, no TIFF output", time.perf_counter() - t0, None))

print("| step | GPU path (s) | CPU oracle (s) |\n|---|---:|---:|")
for name, a, b in rows:
    print(f"| {name} | {a:.3f} | {'' if b is None else format(b, '.2f')} |")
