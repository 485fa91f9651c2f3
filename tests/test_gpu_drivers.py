"""The reference's two callers of the path, mirrored in drivers.py: SimulateMultiViewDataset.main() and the
tile-stitching pair generator (two threads, one context each, SNR sweeps over one convolved volume)."""
import os

import numpy as np
import pytest

from helpers import gaussian_psf, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import mvsim_b200
    return mvsim_b200


def test_main_loop_reduced_size(mv, oracle, tmp_path):
    """main() (:524-665) at 121^3 with 3 views: every stage against the oracle chain, files readable again."""
    psf = gaussian_psf((15, 9, 9), (3.0, 1.2, 1.1), threshold=1e-3)
    r = mv.run_main(str(tmp_path), size=121, angle_increment=120, poissonSNR=-1.0, psf=psf, log=lambda *_: None,
                    keep=("acq", "view", "weights", "con", "psf"))
    assert r["angles"] == [0, 120, 240]
    gt, n = oracle.simulate_phantom(size=121, seed=464232194)
    assert n > 0 and np.array_equal(r["rendered"], gt)
    assert rel_err(r["groundtruth"], oracle.rotate(gt, 0, 15)) <= 1e-6
    for i, angle in enumerate(r["angles"]):
        ref_acq, _, _ = oracle.simulate_view(gt, psf, degrees=angle + 15, inc=3, snr=-1.0, use_fft=True)
        assert r["acq"][i].shape == (41, 121, 121)
        assert rel_err(r["acq"][i], ref_acq) <= 1e-4
        ref_view = oracle.rotate(oracle.make_isotropic(ref_acq, 3), 0, -angle)
        assert rel_err(r["view"][i], ref_view) <= 1e-4
        assert abs(float(r["con"][i].astype(np.float64).mean()) - 1.0) < 1e-4          # adjustImage: mean = avgIntensity
    ws = [oracle.rotate(oracle.weight_image((121, 121, 121)), 0, -a) for a in r["angles"]]
    s_ref = oracle.normalize_weights(ws, 3.0)
    for a, b in zip(r["weights"], ws):
        assert rel_err(a, b) <= 1e-6
    assert rel_err(r["sum_weights"], s_ref) <= 1e-6
    names = sorted(os.listdir(tmp_path))
    for angle in (0, 120, 240):
        for stem in ("rot_view_", "att_view_", "con_view_", "acq_view_", "iso_view_", "aligned_view_", "aligned_view_psf_", "aligned_view_weights"):
            assert f"{stem}{angle}.tif" in names
    assert {"rendered.tif", "groundtruth.tif", "sum_weights.tif"} <= set(names)
    back = mv.tiff.read_float_stack(os.path.join(tmp_path, "acq_view_120.tif"))
    assert np.array_equal(back, r["acq"][1])


def test_main_loop_with_noise_is_reproducible_and_poisson(mv):
    psf = gaussian_psf((9, 7, 7), (2.0, 1.0, 1.0), threshold=1e-3)
    a = mv.run_main(None, size=110, angle_increment=180, poissonSNR=25.0, psf=psf, log=lambda *_: None, keep=("acq", "con"))
    b = mv.run_main(None, size=110, angle_increment=180, poissonSNR=25.0, psf=psf, log=lambda *_: None, keep=("acq",))
    for x, y in zip(a["acq"], b["acq"]):
        assert np.array_equal(x, y)                                  # same seed -> same counts
        assert np.array_equal(x, np.floor(x)) and x.min() >= 0
    lam = a["con"][0][::3].astype(np.float64) * 125.0
    m = lam > 20
    z = (a["acq"][0][m] - lam[m]) / np.sqrt(lam[m])
    assert m.sum() > 1000 and abs(z.mean()) < 0.1 and abs(z.std() - 1) < 0.1
    assert not np.array_equal(a["acq"][0], a["acq"][1])


def test_tile_stitching_pairs(mv, oracle):
    """Two pipelines from two threads, then SNR sweeps re-using the convolved volumes (S/SimulateTileStitching.java)."""
    psf = gaussian_psf((9, 7, 7), (2.0, 1.0, 1.0), threshold=1e-3)
    sts = mv.SimulateTileStitching(None, True, (0.2, 0.2, 0.2), psf=psf, size=110)
    seed = oracle.JavaRandom(464232194).next_int()
    for half, con in ((False, sts.con), (True, sts.conHalfPixel)):
        gt, _ = oracle.simulate_phantom(size=110, half_pixel=half, seed=seed)
        ref = oracle.convolve(oracle.attenuate(gt, 0.01), psf.copy(), "fft")
        oracle.adjust(ref, 0.0001, 1.0)
        assert rel_err(con, ref) <= 1e-4
    assert sts.overlap == [11, 11, 11]
    assert sts.getInterval(0) == ([0, 0, 0], [66, 66, 66]) and sts.getInterval(1) == ([44, 44, 44], [109, 109, 109])
    assert sts.getCorrectTranslation() == [43.5, 43.5, 44 / 3]
    t0, t1 = sts.getNextPair(-1.0)                                    # no noise: exact crops of every 3rd slice
    assert np.array_equal(t0, sts.con[0:67:3, 0:67, 0:67])
    assert np.array_equal(t1, sts.conHalfPixel[44:110:3, 44:110, 44:110])
    means = []
    for snr in (1.0, 4.0, 16.0, 64.0):                                # SNR sweep on the cached volumes
        n0, n1 = sts.getNextPair(snr)
        assert n0.shape == t0.shape and n1.shape == t1.shape and np.array_equal(n0, np.floor(n0))
        means.append(float(n0.mean()) / (snr * snr / 5.0))
    assert np.allclose(means, float(t0.mean()), rtol=0.05)
