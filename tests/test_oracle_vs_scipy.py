"""Independent second opinions on the oracle's semantics (the reference itself cannot run here): scipy.ndimage
implements the same published conventions -- 'mirror' boundary = imglib2 Views.extendMirrorSingle (edge sample not
repeated), order-1 affine resampling with 'grid-constant' = n-linear interpolation over Views.extendZero."""
import math

import numpy as np
import pytest
from scipy import ndimage

from helpers import rel_err


@pytest.mark.parametrize("kshape", [(5, 7, 9), (3, 3, 3), (1, 5, 1)])
def test_convolution_matches_scipy_mirror_boundary(oracle, kshape):
    rng = np.random.default_rng(41)
    vol = rng.random((11, 13, 16), dtype=np.float32)
    psf = rng.random(kshape, dtype=np.float32)
    ours = oracle.convolve(vol, psf, "direct")               # normalises psf in place
    ref = ndimage.convolve(vol.astype(np.float64), psf.astype(np.float64), mode="mirror")    # odd kernels: centre k//2
    assert rel_err(ours, ref) < 1e-6


@pytest.mark.parametrize("degrees", [0, 15, 52, 90, 200, -30])
@pytest.mark.parametrize("shape", [(15, 15, 15), (10, 21, 8)])
def test_rotation_matches_scipy_linear_resampling(oracle, shape, degrees):
    rng = np.random.default_rng(42)
    vol = rng.random(shape, dtype=np.float32)
    ours = oracle.rotate(vol, 0, degrees)
    # inverse map of axisRotation about x: (z, y) of the output -> source; centre (dim-1)//2, float-rounded angle
    th = float(np.float32(math.radians(degrees)))
    c, s = math.cos(th), math.sin(th)
    cz, cy = (shape[0] - 1) // 2, (shape[1] - 1) // 2
    # numpy axis order (z, y, x):  z_src = -s (y-cy) + c (z-cz) + cz ; y_src = c (y-cy) + s (z-cz) + cy
    m = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    off = np.array([cz, cy, 0.0]) - m @ np.array([cz, cy, 0.0])
    ref = ndimage.affine_transform(vol.astype(np.float64), m, offset=off, order=1, mode="grid-constant", cval=0.0)
    assert rel_err(ours, ref) < 2e-6


def test_attenuation_matches_cumulative_product(oracle):
    rng = np.random.default_rng(43)
    vol = rng.random((4, 12, 12), dtype=np.float32) * 2
    ours = oracle.attenuate(vol, 0.05)
    f = np.maximum(1 - 0.05 * vol.astype(np.float64), 0)
    n = np.cumprod(f[:, ::-1, :], axis=1)[:, ::-1, :]        # inclusive product from the top (y = Y-1) down
    assert rel_err(ours, vol * n) < 1e-6


def test_poisson_replay_matches_scipy_distribution(oracle):
    from scipy import stats
    a = np.full(50000, 2.0, dtype=np.float32)
    oracle.poisson(a, 5.0, oracle.JavaRandom(11))            # lambda = 2 * 25/5 = 10
    emp = np.bincount(a.astype(int), minlength=40)[:40] / a.size
    assert np.abs(emp - stats.poisson.pmf(np.arange(40), 10.0)).max() < 6e-3
