"""Host link probe for the e2e arm: H2D alone, D2H alone, both at once (pinned memory, two streams), and the timeline
of two caller threads through mvsim_simulate_views.  Run on the GPU box:  python tools/pcie_probe.py"""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def link():
    n = 1 << 29        # 2 GiB of float32
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    d_a = torch.empty(n, dtype=torch.float32, device="cuda")
    d_b = torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = n * 4 / 1e9

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def up():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        up()
        down()
    t_up, t_down, t_both = timed(up), timed(down), timed(both)
    print(f"H2D alone {gb / t_up:.1f} GB/s, D2H alone {gb / t_down:.1f} GB/s, both at once {gb / t_both:.1f} GB/s each "
          f"({2 * gb / t_both:.1f} GB/s total)")


def callers(n_callers=2, steps=3):
    import bench
    import mvsim_b200 as mv
    shape, kshape, sigma, degrees, inc, snr = bench.WORKLOADS["cfg3"]
    nv = len(degrees)
    oshape = ((shape[0] - 1) // inc + 1, shape[1], shape[2])
    gt = mv.PinnedBuffer(shape)
    gt.array[...] = bench.make_ground_truth(shape)
    psf_raw = bench.make_psfs(kshape, sigma, nv)
    S = mv.SimulateMultiViewDataset
    sets = [(mv.Context(0), [mv.PinnedBuffer(kshape) for _ in range(nv)], [mv.PinnedBuffer(oshape) for _ in range(nv)]) for _ in range(n_callers)]
    log = []

    def step(i):
        c, psfs, outs = sets[i]
        for v in range(nv):
            psfs[v].array[...] = psf_raw[v]
        t0 = time.perf_counter()
        S.simulateViews(gt.array, [p.array for p in psfs], degrees, inc=inc, poissonSNR=snr, rnd=1, ctx=c, outs=[o.array for o in outs])
        log.append((i, t0, time.perf_counter()))
    for i in range(n_callers):
        step(i)
    log.clear()
    th = [threading.Thread(target=lambda i=i: [step(i) for _ in range(steps)]) for i in range(n_callers)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    total = time.perf_counter() - t0
    print(f"{n_callers} callers x {steps} steps: {total * 1e3 / (n_callers * steps):.1f} ms/step")
    for i, a, b in sorted(log, key=lambda r: r[1]):
        print(f"  caller {i}: {1e3 * (a - t0):7.1f} -> {1e3 * (b - t0):7.1f} ms")
    for c, _, _ in sets:
        c.profile(True)
    step(0)
    print({k: round(v[0], 2) for k, v in sets[0][0].stage_times().items() if v[1]})


if __name__ == "__main__":
    print("CUDA_DEVICE_MAX_CONNECTIONS =", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
    if "--link" in sys.argv:
        link()
    callers(2, 3)
