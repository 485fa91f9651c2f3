#!/usr/bin/env python3
"""Profiling driver: one view of BASELINE config 3 on device-resident volumes (no large host arrays), so
`ncu` captures each kernel of the pipeline once.  Usage: python tools/prof_conv.py [views] [workload]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import mvsim_b200 as mv  # noqa: E402
from mvsim_b200._lib import check  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shape, kshape, sigma, degrees, inc, snr = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "cfg3"]
ctx = mv.Context(0)
oshape = ((shape[0] - 1) // inc + 1, shape[1], shape[2])
gt = mv.DeviceVolume(ctx, shape)
# cheap non-trivial content: upload one slab repeatedly would need host memory; use a small phantom tiled by the library's own rotate
small = bench.make_ground_truth((shape[0] // 8, shape[1], shape[2]))
host = np.concatenate([small] * 8, axis=0)
gt.upload(host)
del host, small
psf = mv.DeviceVolume(ctx, kshape, bench.make_psfs(kshape, sigma, 1)[0])
out = mv.DeviceVolume(ctx, oshape)
for v in range(views):
    p = mv.make_view_params(shape, kshape, 0, degrees[v % len(degrees)], 0.01, 0.0001, 1.0, inc, snr, seed=1, stream=v)
    check(ctx._lib.mvsim_dev_simulate_view(ctx.h, C.byref(p), gt.h, psf.h, out.h), ctx.h)
ctx.synchronize()
o = out.download()
print("ok", float(o.mean()), ctx.kernel_launches)
