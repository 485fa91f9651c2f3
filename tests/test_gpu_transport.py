"""uint16 count transport of the batch call (MVSIM_OPT_COUNT_TRANSPORT): the caller still receives the float32 volumes of the
reference's API (extractSlices returns Img<FloatType>, S/SimulateMultiViewDataset.java:195-231), bit-identical to the float32
transport, while half the bytes cross the host link."""
import numpy as np
import pytest

from helpers import gaussian_psf, sphere_phantom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import mvsim_b200
    return mvsim_b200


@pytest.mark.parametrize("threads", [0, 1, 3])
def test_uint16_transport_is_bit_identical(mv, threads):
    S = mv.SimulateMultiViewDataset
    gt = sphere_phantom((45, 72, 68), n_spheres=120)          # plane size not a multiple of 64: ragged chunks for the widening threads
    psfs = [gaussian_psf((11, 9, 7), (2.0 + 0.2 * v, 1.2, 1.0), threshold=1e-3) for v in range(4)]
    degrees = [15, 105, 195, 285]
    plain, fast = mv.Context(0), mv.Context(0).count_transport(True, host_threads=threads)
    for snr in (25.0, 4.0):
        pa, pb = [p.copy() for p in psfs], [p.copy() for p in psfs]
        a = S.simulateViews(gt, pa, degrees, inc=3, poissonSNR=snr, rnd=5, ctx=plain)
        b = S.simulateViews(gt, pb, degrees, inc=3, poissonSNR=snr, rnd=5, ctx=fast)
        for x, y, p, q in zip(a, b, pa, pb):
            assert x.dtype == y.dtype == np.float32 and np.array_equal(x, y)
            assert np.array_equal(p, q)                        # PSFs still normalised in place
    # a device-resident ground truth goes through the same path
    vol = mv.DeviceVolume(fast, gt.shape, gt)
    c = S.simulateViews(vol, [p.copy() for p in psfs], degrees, inc=3, poissonSNR=4.0, rnd=5, ctx=fast)
    for x, y in zip(a, c):
        assert np.array_equal(x, y)
    vol.free()
    plain.close()
    fast.close()


def test_noise_free_views_and_counts_beyond_uint16_fall_back_to_float32(mv):
    S = mv.SimulateMultiViewDataset
    gt = sphere_phantom((24, 40, 40), n_spheres=60)
    psfs = [gaussian_psf((7, 7, 7), (1.5, 1.1, 1.0), threshold=1e-3) for _ in range(2)]
    plain, fast = mv.Context(0), mv.Context(0).count_transport(True)
    # snr < 0: no counts, intensities travel as float32
    a = S.simulateViews(gt, [p.copy() for p in psfs], [15, 195], inc=2, poissonSNR=-1.0, ctx=plain)
    b = S.simulateViews(gt, [p.copy() for p in psfs], [15, 195], inc=2, poissonSNR=-1.0, ctx=fast)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[0].max() < 100 and (a[0] != np.round(a[0])).any()
    # SNR 2000: lambda ~ 8e5 per unit intensity -> counts far beyond 65535 -> the overflow flag sends the view as float32
    a = S.simulateViews(gt, [p.copy() for p in psfs], [15, 195], inc=2, poissonSNR=2000.0, rnd=3, ctx=plain)
    b = S.simulateViews(gt, [p.copy() for p in psfs], [15, 195], inc=2, poissonSNR=2000.0, rnd=3, ctx=fast)
    assert a[0].max() > 65535 and all(np.array_equal(x, y) for x, y in zip(a, b))
    with pytest.raises(mv.MvsimError):
        fast._lib.mvsim_ctx_set_option.restype                  # (binding exists)
        mv._lib.check(fast._lib.mvsim_ctx_set_option(fast.h, 99, 1), fast.h)
    plain.close()
    fast.close()
