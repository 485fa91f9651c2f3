"""Largest-single-volume path on the GPU: SlabConvolution against the plain convolution (one GPU), overlap-save
y blocks at real size, and -- when the box has >= 2 GPUs -- the NCCL all-to-all run under torchrun."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import gaussian_psf, rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mv():
    import mvsim_b200
    return mvsim_b200


def _slab_single(mv, vol, psf):
    import torch
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = mv.Context(0, cuda_stream=stream.cuda_stream)
        sc = mv.SlabConvolution(ctx, vol.shape, psf.shape)
        img = torch.from_numpy(vol).cuda()
        k = torch.from_numpy(psf).cuda()
        out = torch.empty_like(img)
        sc.convolve(img, k, out)
        stream.synchronize()
        res = out.cpu().numpy()
        nb = sc.y_blocks
        sc.close()
        ctx.close()
    return res, nb


def test_slab_plan_on_one_gpu_equals_plain_convolution(mv, oracle):
    rng = np.random.default_rng(21)
    vol = rng.random((24, 40, 56), dtype=np.float32)
    psf = rng.random((5, 7, 9), dtype=np.float32)
    oracle.norm_image(psf)
    got, nb = _slab_single(mv, vol, psf)
    ref = mv.SimulateMultiViewDataset.convolve(vol, psf.copy())
    assert nb == 1 and np.array_equal(got, ref)
    assert rel_err(got, oracle.convolve(vol, psf.copy(), "direct")) <= 1e-5


def test_y_lines_beyond_the_size_table_use_overlap_save_blocks(mv):
    # Y + KY - 1 = 1720 > 1600: two blocks; sparse kernel so the expected value is a 3-tap sum
    rng = np.random.default_rng(22)
    vol = rng.random((6, 1700, 24), dtype=np.float32)
    psf = np.zeros((3, 21, 5), dtype=np.float32)
    taps = [((0, 0, 0), 0.5), ((2, 20, 4), 0.3), ((1, 10, 2), 0.2)]
    for t, w in taps:
        psf[t] = w
    got, nb = _slab_single(mv, vol, psf)
    assert nb == 2
    plain = mv.SimulateMultiViewDataset.convolve(vol, psf.copy())
    assert rel_err(plain, got) <= 1e-6

    def mir(i, n):
        p = 2 * (n - 1); j = i % p
        return p - j if j >= n else j
    for (z, y, x) in [(0, 0, 0), (5, 1699, 23), (3, 849, 11), (3, 850, 11), (3, 851, 12), (2, 1690, 1)]:
        exp = sum(w * vol[mir(z - (t[0] - 1), 6), mir(y - (t[1] - 10), 1700), mir(x - (t[2] - 2), 24)] for t, w in taps)
        assert got[z, y, x] == pytest.approx(exp, rel=2e-5)


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_multi_gpu_slab_convolution_matches_single_gpu(tmp_path, mode):
    """p2p: exchanges fused into the kernels as NVLink peer stores; nccl: two all_to_all_single per y block."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    out = tmp_path / "slab.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(HERE, "slab_worker.py"), "32x72x118", "9x7x11", str(out), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(out.read_text())
    assert d["world"] == world and d["p2p"] == (mode == "p2p") and d["identical"], d
