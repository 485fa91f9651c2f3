"""Worker of tests/test_gpu_slab.py: ONE view of a z-slab-decomposed volume (SlabView) on `world` ranks -- sharing GPU 0 ("shared":
CUDA IPC + gloo) or one GPU each (NCCL) -- compared on rank 0 with the undecomposed mvsim_simulate_view of the same inputs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import mvsim_b200 as mv  # noqa: E402
from helpers import gaussian_psf, sphere_phantom  # noqa: E402


def main():
    shape = tuple(int(v) for v in sys.argv[1].split("x"))       # Z x Y x X
    kshape = tuple(int(v) for v in sys.argv[2].split("x"))
    out_path, inc, shared = sys.argv[3], int(sys.argv[4]), sys.argv[5] == "shared"
    local = 0 if shared else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    grp = mv.Group("gloo" if shared else "nccl", device=dev)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = mv.Context(local, cuda_stream=stream.cuda_stream)
    gt = sphere_phantom(shape, n_spheres=150)
    psf = gaussian_psf(kshape, (kshape[0] / 6.0, kshape[1] / 7.0, kshape[2] / 8.0), threshold=1e-3)
    sv = mv.SlabView(ctx, shape, kshape, grp.rank, grp.world, grp.dist, p2p=True)
    d_gt = torch.from_numpy(gt).to(dev)
    res = {}
    for snr in (-1.0, 25.0):
        d_psf = torch.from_numpy(psf.copy()).to(dev)
        kept, con = sv.simulate(d_gt, d_psf, degrees=75, inc=inc, snr=snr, seed=77, stream=3)
        torch.cuda.synchronize()
        kept, con = kept.cpu(), con.cpu()
        k0, nk = sv.kept_planes(inc)
        rec = {"k0": k0, "kept": kept.numpy(), "con": con.numpy(), "z0": sv.z0}
        allrec = [None] * grp.world if grp.rank == 0 else None
        if grp.world > 1:
            grp.dist.gather_object(rec, allrec, dst=0)
        else:
            allrec = [rec]
        if grp.rank == 0:
            allrec.sort(key=lambda r: r["z0"])
            got = np.concatenate([r["kept"] for r in allrec if r["kept"].shape[0]], axis=0)
            got_con = np.concatenate([r["con"] for r in allrec], axis=0)
            ref = mv.SimulateMultiViewDataset.simulateView(gt, psf.copy(), 75, inc=inc, poissonSNR=snr, rnd=77, ctx=ctx, stream=3)
            S = mv.SimulateMultiViewDataset
            con_ref = S.convolve(S.attenuate3d(S.rotateAroundAxis(gt, 0, 75, ctx=ctx), 0.01, ctx=ctx), psf.copy(), ctx=ctx)
            mv.Tools.adjustImage(con_ref, 0.0001, 1.0, ctx=ctx)
            key = "clean" if snr < 0 else "noisy"
            res[key] = {"shape_ok": got.shape == ref.shape,
                        "max_rel_err": float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max()),
                        "identical_fraction": float((got == ref).mean()),
                        "conv_max_rel_err": float(np.abs(got_con.astype(np.float64) - con_ref).max() / np.abs(con_ref).max())}
    if grp.rank == 0:
        with open(out_path, "w") as f:
            json.dump({"world": grp.world, "shared_gpu": shared, **res}, f)
    sv.close()
    grp.close()


if __name__ == "__main__":
    main()
