"""Size-independent properties at BASELINE.json's full size (config 3: 1024x1024x512, PSF 128^3),
where the CPU oracle cannot follow: constant image, impulse response, slice selection."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE = (512, 1024, 1024)      # (Z, Y, X)
KSHAPE = (128, 128, 128)


@pytest.fixture(scope="module")
def mv():
    import mvsim_b200
    return mvsim_b200


def _psf():
    rng = np.random.default_rng(12)
    psf = np.zeros(KSHAPE, dtype=np.float32)
    idx = rng.integers(0, 128, size=(400, 3))
    psf[idx[:, 0], idx[:, 1], idx[:, 2]] = rng.random(400, dtype=np.float32) + 0.1
    psf[0, 0, 0] = 1.0
    psf[127, 127, 127] = 2.0
    return psf


def test_fullsize_constant_image_and_impulse(mv):
    S = mv.SimulateMultiViewDataset
    ctx = mv.Context(0)
    psf = _psf()
    vol = np.full(SHAPE, 1.5, dtype=np.float32)
    out = S.convolve(vol, psf, ctx=ctx)                 # psf normalised in place
    assert float(psf.astype(np.float64).sum()) == pytest.approx(1.0, abs=1e-5)
    assert np.abs(out - 1.5).max() <= 1.5e-4 * 1.5      # sum(psf)=1 + mirror border => constant
    vol[:] = 0
    p = (200, 700, 300)
    vol[p] = 1.0
    out = S.convolve(vol, psf, ctx=ctx)
    # out[x] = psfN[x - p + kdim/2]: un-flipped PSF centred on the impulse
    win = out[p[0] - 64:p[0] + 64, p[1] - 64:p[1] + 64, p[2] - 64:p[2] + 64]
    assert np.abs(win - psf).max() <= 1e-4 * psf.max()
    out[p[0] - 64:p[0] + 64, p[1] - 64:p[1] + 64, p[2] - 64:p[2] + 64] = 0
    assert np.abs(out).max() <= 1e-4 * psf.max()
    ctx.close()


def test_fullsize_view_slice_selection_and_mean(mv):
    S = mv.SimulateMultiViewDataset
    ctx = mv.Context(0)
    rng = np.random.default_rng(13)
    gt = np.zeros(SHAPE, dtype=np.float32)
    gt[128:384, 256:768, 256:768] = rng.random((256, 512, 512), dtype=np.float32)
    psf = np.zeros(KSHAPE, dtype=np.float32)
    psf[60:68, 60:68, 60:68] = 1.0
    acq = S.simulateView(gt, psf, 75, inc=5, poissonSNR=-1.0, ctx=ctx)
    assert acq.shape == (103, 1024, 1024)
    # adjustImage: mean over the whole convolved volume is 1; every 5th slice is an unbiased sample of it
    # background: minValue +- float32 FFT round-off (the reference's FFT has the same ~1e-7 * max noise)
    assert acq.min() >= 1e-4 - 1e-6 * float(acq.max())
    assert 0.8 < float(acq.astype(np.float64).mean()) < 1.25
    noisy = S.simulateView(gt, psf, 75, inc=5, poissonSNR=25.0, rnd=3, ctx=ctx)
    lam = acq.astype(np.float64) * 125.0
    z = (noisy - lam) / np.sqrt(lam)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    ctx.close()
