"""Worker of tests/test_gpu_slab.py: one process per GPU under torchrun (NCCL).  Every rank convolves its z slab
of a seeded global volume with SlabConvolution (two NCCL all-to-all transposes); rank 0 compares the gathered
result with the undecomposed single-GPU convolution of the same volume."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import mvsim_b200 as mv  # noqa: E402
from helpers import gaussian_psf  # noqa: E402


def main():
    shape = tuple(int(v) for v in sys.argv[1].split("x"))       # Z x Y x X
    kshape = tuple(int(v) for v in sys.argv[2].split("x"))
    out_path = sys.argv[3]
    p2p = len(sys.argv) < 5 or sys.argv[4] != "nccl"
    # "shared": every rank runs on GPU 0 (the driver's 1-GPU box): CUDA IPC maps the peers' buffers between the processes,
    # gloo carries the handles, the barriers and (non-p2p mode) the all-to-all -- NCCL refuses two ranks on one device
    shared = len(sys.argv) > 5 and sys.argv[5] == "shared"
    local = 0 if shared else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    grp = mv.Group("gloo" if shared else "nccl", device=dev)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = mv.Context(local, cuda_stream=stream.cuda_stream)
    rng = np.random.default_rng(2024)
    vol = rng.random(shape, dtype=np.float32)
    psf = gaussian_psf(kshape, (kshape[0] / 7.0, kshape[1] / 6.0, kshape[2] / 5.0))
    mv.Tools.normImage(psf, ctx=ctx)
    sc = mv.SlabConvolution(ctx, shape, kshape, grp.rank, grp.world, grp.dist, p2p=p2p)
    z0, zl = sc.z0, sc.z_local
    img = torch.from_numpy(vol[z0:z0 + zl]).to(dev)
    d_psf = torch.from_numpy(psf).to(dev)
    out = torch.empty_like(img)
    sc.convolve(img, d_psf, out)
    torch.cuda.synchronize()
    if shared:
        out = out.cpu()
    parts = [torch.empty_like(out) for _ in range(grp.world)] if grp.rank == 0 else None
    if grp.world > 1:
        grp.dist.gather(out, parts, dst=0)
    else:
        parts = [out]
    if grp.rank == 0:
        got = torch.cat(parts, dim=0).cpu().numpy()
        ref = mv.SimulateMultiViewDataset.convolve(vol, psf.copy(), ctx=ctx)      # undecomposed, same GPU kernels
        err = float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max())
        with open(out_path, "w") as f:
            json.dump({"world": grp.world, "p2p": sc.p2p, "shared_gpu": shared, "y_blocks": sc.y_blocks, "nfft": sc.nfft, "max_rel_err": err,
                       "identical": bool(np.array_equal(got, ref))}, f)
    sc.close()
    grp.close()


if __name__ == "__main__":
    main()
