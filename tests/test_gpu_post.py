"""Post-acquisition chain of main(): makeIsotropic, computeWeightImage, rotate-back and the
cross-view weight normalisation, against the oracle."""
import numpy as np
import pytest

from helpers import rel_err, sphere_phantom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import mvsim_b200
    return mvsim_b200.SimulateMultiViewDataset


@pytest.mark.parametrize("shape,inc", [((7, 9, 12), 3), ((5, 8, 8), 5), ((1, 4, 4), 3), ((10, 6, 7), 1)])
def test_make_isotropic_bit_exact(S, oracle, shape, inc):
    v = np.random.default_rng(31).random(shape, dtype=np.float32)
    got = S.makeIsotropic(v, inc)
    ref = oracle.make_isotropic(v, inc)
    assert got.shape == ((shape[0] - 1) * inc + 1, shape[1], shape[2])
    assert np.array_equal(got, ref)
    assert np.array_equal(got[::inc], v)            # acquired planes are reproduced exactly


@pytest.mark.parametrize("shape", [(3, 100, 5), (2, 289, 4), (2, 30, 3)])
def test_weight_image_bit_exact(S, oracle, shape):
    got = S.computeWeightImage(shape, 0.01)
    assert np.array_equal(got, oracle.weight_image(shape))
    assert got[0, -1, 0] == 1.0 and (shape[1] < 90 or got[0, 0, 0] == 0.0)


def test_weight_normalisation_matches_oracle(S, oracle):
    rng = np.random.default_rng(32)
    ws = [rng.random((6, 10, 12), dtype=np.float32) * (rng.random((6, 10, 12)) > 0.3) for _ in range(7)]
    ws = [np.ascontiguousarray(w, dtype=np.float32) for w in ws]
    ws[0][0, 0, :] = 0
    for w in ws[1:]:
        w[0, 0, :4] = 0                               # voxels no view covers stay 0
    ref = [w.copy() for w in ws]
    s_ref = oracle.normalize_weights(ref, 3.0)
    s = S.normalizeWeights(ws, 3.0)
    for a, b in zip(ws, ref):
        assert np.array_equal(a, b)
    assert np.array_equal(s, s_ref) and np.all(s[0, 0, :4] == 0) and s.max() <= 7.0


def test_whole_main_loop_for_one_view(S, oracle):
    """acq -> makeIsotropic -> rotate back by -angle, and the weight image rotated back (:585-593)."""
    gt = sphere_phantom((46, 45, 45), n_spheres=300)       # (Z-1) divisible by inc, like 289 = 96*3+1
    angle = 52
    psf = np.zeros((9, 9, 9), dtype=np.float32)
    psf[2:7, 3:6, 3:6] = 1
    acq = S.simulateView(gt, psf.copy(), angle + 15, inc=3, poissonSNR=-1.0)
    iso = S.makeIsotropic(acq, 3)
    assert iso.shape == gt.shape
    view = S.rotateAroundAxis(iso, 0, -angle)
    ref_acq, _, _ = oracle.simulate_view(gt, psf, degrees=angle + 15, inc=3, snr=-1.0, use_fft=False)
    ref_view = oracle.rotate(oracle.make_isotropic(ref_acq, 3), 0, -angle)
    assert rel_err(view, ref_view) <= 1e-4
    w = S.rotateAroundAxis(S.computeWeightImage(gt.shape), 0, -angle)
    assert rel_err(w, oracle.rotate(oracle.weight_image(gt.shape), 0, -angle)) <= 1e-6
