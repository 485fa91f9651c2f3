"""PSF-spectrum cache: repeated PSFs skip the PSF transforms, results stay bit-identical, the
in-place normalisation of the caller's PSF (S/SimulateMultiViewDataset.java:255) stays observable, a changed PSF misses."""
import numpy as np
import pytest

from helpers import gaussian_psf, sphere_phantom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import mvsim_b200
    return mvsim_b200


def test_repeated_psf_hits_changed_psf_misses_results_identical(mv):
    S = mv.SimulateMultiViewDataset
    gt = sphere_phantom((48, 64, 64), n_spheres=80)
    psf = gaussian_psf((13, 9, 9), (2.5, 1.2, 1.1), threshold=1e-3)
    plain = mv.Context(0)                                   # cache off: the library default
    p0 = psf.copy()
    ref = S.convolve(gt, p0, ctx=plain)
    assert plain.psf_cache_stats() == {"hits": 0, "misses": 0, "entries": 0, "bytes": 0}

    ctx = mv.Context(0).psf_cache(256 << 20)
    p1 = psf.copy()
    a = S.convolve(gt, p1, ctx=ctx)                         # miss: spectrum computed into a cache entry
    assert ctx.psf_cache_stats()["misses"] == 1 and ctx.psf_cache_stats()["entries"] == 1
    p2 = psf.copy()
    b = S.convolve(gt, p2, ctx=ctx)                         # hit
    c = S.convolve(gt, p2, ctx=ctx)                         # the caller passes the already-normalised PSF again (tile pairs): hit
    st = ctx.psf_cache_stats()
    assert st["hits"] == 2 and st["misses"] == 1 and st["entries"] == 1
    assert np.array_equal(a, ref) and np.array_equal(b, ref) and np.array_equal(c, ref)
    assert np.array_equal(p1, p0) and np.array_equal(p2, p0)               # normalised in place on every call, hit or miss
    assert abs(float(p2.astype(np.float64).sum()) - 1.0) < 1e-6

    changed = psf.copy()
    changed[6, 4, 4] *= 1.0 + 2.0 ** -20                   # one voxel, a few ulps
    d = S.convolve(gt, changed, ctx=ctx)
    assert ctx.psf_cache_stats()["misses"] == 2 and not np.array_equal(d, ref)
    # same PSF, other volume: another plan (padded sizes), hence another entry
    S.convolve(gt[:40], psf.copy(), ctx=ctx)
    st = ctx.psf_cache_stats()
    assert st["misses"] == 3 and st["entries"] == 3 and st["bytes"] > 0
    # shrinking the budget drops everything; a budget below one entry switches caching off for that plan
    ctx.psf_cache(1024)
    assert ctx.psf_cache_stats()["entries"] == 0
    e = S.convolve(gt, psf.copy(), ctx=ctx)
    assert np.array_equal(e, ref) and ctx.psf_cache_stats()["entries"] == 0
    ctx.close()
    plain.close()


def test_snr_sweep_and_view_loop_skip_the_psf_stage(mv):
    S = mv.SimulateMultiViewDataset
    gt = sphere_phantom((40, 56, 56), n_spheres=60)
    psfs = [gaussian_psf((15, 7, 7), (3.0 + 0.2 * v, 1.0, 0.9), threshold=1e-3) for v in range(3)]
    degrees = [15, 135, 255]
    ctx = mv.Context(0).psf_cache(128 << 20)
    ref_ctx = mv.Context(0)
    for sweep, snr in enumerate((4.0, 25.0, 100.0)):      # one PSF per view, re-used across the SNR sweep
        got = S.simulateViews(gt, [p.copy() for p in psfs], degrees, inc=3, poissonSNR=snr, rnd=9, ctx=ctx)
        ref = S.simulateViews(gt, [p.copy() for p in psfs], degrees, inc=3, poissonSNR=snr, rnd=9, ctx=ref_ctx)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
        st = ctx.psf_cache_stats()
        assert st["misses"] == 3 and st["hits"] == 3 * sweep
    # LRU: a budget of two entries keeps the two most recently used spectra
    per_entry = ctx.psf_cache_stats()["bytes"] // 3
    ctx.psf_cache(0)
    ctx.psf_cache(2 * per_entry + 16)
    for v in (0, 1, 2, 2, 1, 0):
        S.simulateView(gt, psfs[v].copy(), degrees[v], inc=3, poissonSNR=-1.0, ctx=ctx)
    st = ctx.psf_cache_stats()
    assert st["entries"] == 2 and st["misses"] == 3 + 4 and st["hits"] == 6 + 2       # 0 1 2 miss, 2 1 hit, 0 miss (evicted)
    # the profile shows what a hit skips: normalise + hash only, no x / y transform of the PSF
    ctx.profile(True)
    S.simulateView(gt, psfs[0].copy(), degrees[0], inc=3, poissonSNR=-1.0, ctx=ctx)
    hit_launches = ctx.stage_times()["psf"][1]
    ctx.psf_cache(0)
    ctx.profile(True)
    S.simulateView(gt, psfs[0].copy(), degrees[0], inc=3, poissonSNR=-1.0, ctx=ctx)
    cold_launches = ctx.stage_times()["psf"][1]
    assert hit_launches < cold_launches
    ctx.close()
    ref_ctx.close()
