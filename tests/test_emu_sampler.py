"""The Poisson sampler's own code (sampler.cuh, __host__ __device__) evaluated on the CPU: Philox4x32-10
known-answer vectors (Random123 kat_vectors) and distribution checks against scipy's exact pmf and
against the oracle's replay of the reference sampler (S/uncommons/PoissonGenerator.java:95-109)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "multiview-simulation_b200", "csrc")
SO = os.path.join(HERE, "emu", "libemu_sampler.so")
SRC = os.path.join(HERE, "emu", "emu_sampler.cpp")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC, os.path.join(CSRC, "sampler.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fno-gnu-unique", "-Wl,-Bsymbolic", "-I", CSRC, SRC, "-o", SO])
    L = C.CDLL(SO)
    L.emu_poisson.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_uint64, C.c_uint64, C.c_uint64]
    return L


def sample(emu, lam, n, seed=1, stream=0):
    l = np.full(n, lam, dtype=np.float64)
    out = np.empty(n, dtype=np.float32)
    emu.emu_poisson(l.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_float)), n, seed, stream)
    return out


def test_philox_known_answers(emu):
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = (C.c_uint32 * 4)()
        emu.emu_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == exp


@pytest.mark.parametrize("lam", [0.01, 0.5, 3.0, 9.99, 10.0, 30.0, 300.0, 3000.0, 1.0e5])
def test_distribution_matches_exact_pmf(emu, lam):
    from scipy import stats
    n = 400000
    a = sample(emu, lam, n)
    assert np.array_equal(a, np.round(a)) and a.min() >= 0
    assert abs(a.mean() - lam) < 5 * math.sqrt(lam / n)
    assert abs(a.var() - lam) < 6 * lam * math.sqrt(2.0 / n) + 6 * math.sqrt(lam / n)
    # chi-square against the exact pmf on bins with expectation >= 20
    lo, hi = int(stats.poisson.ppf(1e-4, lam)), int(stats.poisson.ppf(1 - 1e-4, lam))
    ks = np.arange(lo, hi + 1)
    step = max(1, len(ks) // 60)
    edges = list(ks[::step]) + [hi + 1]
    obs, exp = [], []
    for e0, e1 in zip(edges[:-1], edges[1:]):
        obs.append(np.sum((a >= e0) & (a < e1)))
        exp.append(n * (stats.poisson.cdf(e1 - 1, lam) - stats.poisson.cdf(e0 - 1, lam)))
    obs, exp = np.array(obs, float), np.array(exp, float)
    keep = exp >= 20
    chi2 = np.sum((obs[keep] - exp[keep]) ** 2 / exp[keep])
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4


@pytest.mark.parametrize("lam", [0.5, 30.0, 750.0])
def test_ks_against_reference_sampler(emu, oracle, lam):
    from scipy import stats
    snr = 10.0
    mul = (snr / math.sqrt(5)) ** 2
    v = np.float32(lam / mul)
    n = 30000
    ref = np.full(n, v, dtype=np.float32)
    oracle.poisson(ref, snr, oracle.JavaRandom(3))
    got = sample(emu, float(v) * mul, 3 * n)
    assert stats.ks_2samp(got, ref).pvalue > 0.01


def test_edge_cases_and_streams(emu):
    assert np.all(sample(emu, 0.0, 100) == 0) and np.all(sample(emu, -3.0, 100) == 0)
    assert np.all(sample(emu, float("nan"), 10) == 0)
    a, b, c = sample(emu, 5.0, 1000, 1, 0), sample(emu, 5.0, 1000, 1, 1), sample(emu, 5.0, 1000, 2, 0)
    assert np.array_equal(a, sample(emu, 5.0, 1000, 1, 0))
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    big = sample(emu, 3.0e8, 20000)
    assert abs(big.mean() / 3.0e8 - 1) < 1e-5 and abs(big.std() / math.sqrt(3.0e8) - 1) < 0.05


def test_uniform_is_strictly_inside_the_unit_interval(emu):
    """with 24 bits, (float)(2^24 - 1) + 0.5f rounded to 2^24 and u01f returned exactly 1.0f."""
    emu.emu_u01f.restype = C.c_float
    emu.emu_u01f.argtypes = [C.c_uint32]
    words = [0, 1, 0x1ff, 0x200, 0x7fffffff, 0x80000000, 0xfffffdff, 0xfffffe00, 0xffffff00, 0xfffffffe, 0xffffffff]
    us = [emu.emu_u01f(w) for w in words]
    assert all(0.0 < u < 1.0 for u in us)
    assert us[0] == 2.0 ** -24 and us[-1] == 1.0 - 2.0 ** -24
    assert us == sorted(us)


@pytest.mark.parametrize("lam", [0.0125, 0.5, 1.0, 5.0, 9.99])
def test_inversion_search_terminates_on_the_cdf_plateau(emu, lam):
    """The float32 CDF saturates below 1; the top random words must not run the search to its iteration cap (they returned
    64 at lam = 0.0125, the background rate at the default SNR 25 / minValue 1e-4)."""
    from scipy import stats
    emu.emu_poisson_fast.restype = C.c_int
    emu.emu_poisson_fast.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    out = C.c_float()
    hi = stats.poisson.ppf(1 - 1e-9, lam) + 3          # a generous bound on any plausible draw
    for w in [0xffffffff, 0xfffffffe, 0xfffffe00, 0xfffffdff, 0xfffffc00, 0xfffff000]:
        assert emu.emu_poisson_fast(lam, w, 0, C.byref(out)) == 1
        assert 0 <= out.value <= hi, (lam, hex(w), out.value)
    # the median word gives the median count
    assert emu.emu_poisson_fast(lam, 0x80000000, 0, C.byref(out)) == 1
    assert out.value == stats.poisson.ppf(0.5, lam)
