"""CUDA path against the committed golden vectors (tests/golden/): every stage on the seeded inputs, and the whole view on
the reference's own PSF fixture (Angle0.tif)."""
import numpy as np
import pytest

from helpers import rel_err, sphere_phantom
from test_golden import GOLDEN, load_psf

pytestmark = pytest.mark.gpu
TOL = 1e-4          # north_star: max relative error <= 1e-4 in float32


@pytest.fixture(scope="module")
def g():
    import os
    return np.load(os.path.join(GOLDEN, "stages_small.npz"))


@pytest.fixture(scope="module")
def mv():
    import mvsim_b200
    return mvsim_b200


def test_stages_against_golden(mv, g):
    S, T = mv.SimulateMultiViewDataset, mv.Tools
    gt = g["gt"]
    assert rel_err(S.rotateAroundAxis(gt, 0, 75), g["rot_75"]) <= 1e-6
    assert rel_err(S.rotateAroundAxis(gt, 1, 33), g["rot_axis1_33"]) <= 1e-6
    assert np.array_equal(S.attenuate3d(g["rot_75"], 0.01), g["att"])
    k = g["psf_crop"].copy()
    con = S.convolve(g["att"], k)
    assert rel_err(k, g["psf_norm"]) <= 1e-6                      # normalised in place (:255)
    assert rel_err(con, g["con"]) <= TOL
    adj = g["con"].copy()
    corr = T.adjustImage(adj, 0.0001, 1.0)
    assert abs(corr - float(g["corr"])) <= 1e-9 * abs(corr) and rel_err(adj, g["adj"]) <= 1e-6
    acq = S.extractSlices(g["adj"], 3, -1.0)
    assert np.array_equal(acq, g["acq_nonoise"])                   # slice selection is bit exact
    assert np.array_equal(S.makeIsotropic(g["acq_nonoise"], 3), g["iso"])
    assert np.array_equal(S.computeWeightImage((6, 100, 5)), g["weight"])


def test_generators_against_golden(mv, g):
    S, B = mv.SimulateMultiViewDataset, mv.SimulateBeads
    interval = B.interval((64, 48, 40))
    pts = B.randomPoints(40, interval, 535)
    assert np.array_equal(pts, g["bead_points"])
    img = B.renderPoints(B.transformPoints(pts, [45], 0, interval), interval, (1.0, 1.0, 3.0))[0]
    assert rel_err(img, g["beads"]) <= 1e-6
    big, n = S.drawSpheres((242, 242, 242), scale=2, rnd=464232194, return_count=True)
    assert n == len(g["sphere_list"])
    assert np.array_equal(S.downSample2x(big), g["phantom_120"])
    assert np.array_equal(mv.Tools.makeSquare(g["square_in"]), g["square_out"])


def test_poisson_statistics_against_golden_replay(mv, g):
    """Same noise-free intensities as the golden java.util.Random replay: per-lambda mean and variance agree (north_star)."""
    ramp = np.repeat(g["poisson_ramp_in"][::64], 4096).astype(np.float32)           # 64 lambda levels x 4096 samples
    noisy = ramp.copy()
    mv.Tools.poissonProcess(noisy, 25.0, 12345)
    lam = ramp.reshape(64, 4096)[:, 0].astype(np.float64) * 125.0
    got = noisy.reshape(64, 4096)
    assert np.all(np.abs(got.mean(1) - lam) <= 5 * np.sqrt(np.maximum(lam, 1e-9) / 4096) + 1e-12)
    # sample variance of n Poisson draws: sd = sqrt((lam + 2 lam^2) / n); 6 sigma per level
    assert np.all(np.abs(got.var(1) - lam) <= 6 * np.sqrt((lam + 2 * lam ** 2) / 4096) + 1e-12)
    ref = g["poisson_ramp_out"].astype(np.float64)
    lam_ref = g["poisson_ramp_in"].astype(np.float64) * 125.0
    m = lam_ref > 50
    assert abs(((ref[m] - lam_ref[m]) / np.sqrt(lam_ref[m])).std() - ((got[lam > 50] - lam[lam > 50, None]) / np.sqrt(lam[lam > 50, None])).std()) < 0.1


def test_whole_view_on_the_reference_psf_fixture(mv, oracle):
    """Config 0 shape of work at reduced size: the reference's measured-style 51^3 PSF, view at 52 + 15 degrees."""
    S = mv.SimulateMultiViewDataset
    psf = load_psf()
    gt = sphere_phantom((61, 64, 64), n_spheres=150)
    acq = S.simulateView(gt, psf.copy(), 67, inc=3, poissonSNR=-1.0)
    ref, _, _ = oracle.simulate_view(gt, psf, degrees=67, inc=3, snr=-1.0, use_fft=True)
    assert acq.shape == (21, 64, 64) and rel_err(acq, ref) <= TOL
