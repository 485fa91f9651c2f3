#!/usr/bin/env python3
"""Quick A/B timer: N views of a workload on device-resident volumes, CUDA-event stage times per view.
Usage: python tools/time_view.py [reps] [workload]   (environment knobs such as MVSIM_XY_CHUNK are read by the library)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import mvsim_b200 as mv  # noqa: E402
from mvsim_b200._lib import check  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shape, kshape, sigma, degrees, inc, snr = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "cfg3"]
ctx = mv.Context(0)
oshape = ((shape[0] - 1) // inc + 1, shape[1], shape[2])
gt = mv.DeviceVolume(ctx, shape)
small = bench.make_ground_truth((shape[0] // 8, shape[1], shape[2]))
gt.upload(np.concatenate([small] * 8, axis=0))
psf_host = bench.make_psfs(kshape, sigma, 1)[0]
psf = mv.DeviceVolume(ctx, kshape, psf_host)
out = mv.DeviceVolume(ctx, oshape)


def run(n):
    for v in range(n):
        p = mv.make_view_params(shape, kshape, 0, degrees[v % len(degrees)], 0.01, 0.0001, 1.0, inc, snr, seed=1, stream=v)
        check(ctx._lib.mvsim_dev_simulate_view(ctx.h, C.byref(p), gt.h, psf.h, out.h), ctx.h)
    ctx.synchronize()


run(3)
ctx.profile(True)
t0 = time.perf_counter()
run(reps)
wall = (time.perf_counter() - t0) / reps * 1e3
st = ctx.stage_times()
o = out.download()
knobs = {k: v for k, v in os.environ.items() if k.startswith("MVSIM_")}
print(f"{knobs} wall {wall:.3f} ms/view  sum {sum(v[0] for v in st.values()) / reps:.3f}  " +
      " ".join(f"{k}={v[0] / reps:.3f}" for k, v in st.items() if v[1]) + f"  checksum {float(o[::5, ::37, ::41].astype(np.float64).mean()):.9f}")
