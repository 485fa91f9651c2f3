"""ctypes loader for the CPU oracle (oracle/mvsim_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.  PARITY UNPINNED (see the header
of mvsim_oracle.c): the reference has no tests or golden outputs and cannot run here.

Volumes are numpy float32 arrays of shape (Z, Y, X), C-contiguous, i.e. the reference's
ArrayImg order (x fastest); `dims` passed to C are (X, Y, Z).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvsim_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "mvsim_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64p = C.POINTER(C.c_int64)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        L.orc_axis_rotation.argtypes = [i64p, C.c_int, C.c_int, dp]
        L.orc_affine_invert.argtypes = [dp, dp]
        L.orc_rotate.argtypes = [fp, fp, i64p, C.c_int, C.c_int]
        L.orc_attenuate.argtypes = [fp, fp, i64p, C.c_double, C.c_int]
        L.orc_sum.argtypes = [fp, C.c_size_t]
        L.orc_sum.restype = C.c_double
        L.orc_norm_image.argtypes = [fp, C.c_size_t]
        L.orc_norm_image.restype = C.c_double
        L.orc_convolve_direct.argtypes = [fp, i64p, fp, i64p, fp]
        L.orc_convolve_fft.argtypes = [fp, i64p, fp, i64p, fp, C.c_int]
        L.orc_adjust.argtypes = [fp, C.c_size_t, C.c_float, C.c_float]
        L.orc_adjust.restype = C.c_double
        L.orc_extract_slices.argtypes = [fp, i64p, C.c_int, C.c_float, C.c_int64, fp]
        L.orc_poisson.argtypes = [fp, C.c_size_t, C.c_double, C.c_void_p]
        L.orc_poisson.restype = None
        L.orc_jrandom_init.argtypes = [C.c_void_p, C.c_int64]
        L.orc_jrandom_next_int.argtypes = [C.c_void_p]
        L.orc_jrandom_next_int.restype = C.c_int32
        L.orc_jrandom_next_long.argtypes = [C.c_void_p]
        L.orc_jrandom_next_long.restype = C.c_int64
        L.orc_jrandom_next_double.argtypes = [C.c_void_p]
        L.orc_jrandom_next_double.restype = C.c_double
        L.orc_fft_size.argtypes = [C.c_int64]
        L.orc_fft_size.restype = C.c_int64
        L.orc_set_stage_threads.argtypes = [C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_hw_threads.restype = C.c_int
        L.orc_set_poisson_threads.argtypes = [C.c_int]
        L.orc_make_isotropic.argtypes = [fp, i64p, C.c_int, fp]
        L.orc_weight_image.argtypes = [i64p, fp]
        L.orc_normalize_weights.argtypes = [C.POINTER(fp), C.c_int, C.c_size_t, C.c_float, fp]
        L.orc_random_points.argtypes = [C.c_int, i64p, C.c_int64, dp]
        L.orc_random_points.restype = None
        L.orc_transform_points.argtypes = [dp, C.c_int, i64p, C.c_int, C.c_int, dp]
        L.orc_render_beads.argtypes = [dp, C.c_int, dp, i64p, i64p, fp]
        L.orc_draw_spheres.argtypes = [fp, i64p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int64, fp, C.c_int64]
        L.orc_draw_spheres.restype = C.c_int64
        L.orc_downsample2x.argtypes = [fp, i64p, fp]
        L.orc_simulate_phantom.argtypes = [C.c_int, C.c_int, C.c_int64, fp]
        L.orc_simulate_phantom.restype = C.c_int64
        L.orc_make_square.argtypes = [fp, i64p, fp]
        L.orc_jrandom_next_int_bound.argtypes = [C.c_void_p, C.c_int32]
        L.orc_jrandom_next_int_bound.restype = C.c_int32
        L.orc_simulate_view.argtypes = [fp, i64p, fp, i64p, C.c_int, C.c_int, C.c_double, C.c_float,
                                        C.c_float, C.c_int, C.c_float, C.c_int64, C.c_int, C.c_int,
                                        fp, fp, dp]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dims(a):
    z, y, x = a.shape
    return (C.c_int64 * 3)(x, y, z)


def _vol(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 3
    return a


def _check(err, what):
    if err != 0:
        raise ValueError(f"oracle {what}: error {err}")


class JavaRandom:
    """java.util.Random, bit-exact (JDK spec)."""

    def __init__(self, seed):
        self._s = C.c_uint64(0)
        lib().orc_jrandom_init(C.byref(self._s), int(seed))

    def next_int(self):
        return lib().orc_jrandom_next_int(C.byref(self._s))

    def next_long(self):
        return lib().orc_jrandom_next_long(C.byref(self._s))

    def next_double(self):
        return lib().orc_jrandom_next_double(C.byref(self._s))

    def next_int_bound(self, bound):
        return lib().orc_jrandom_next_int_bound(C.byref(self._s), int(bound))

    @property
    def ptr(self):
        return C.byref(self._s)


def axis_rotation(dims_xyz, axis, degrees):
    m = (C.c_double * 12)()
    _check(lib().orc_axis_rotation((C.c_int64 * 3)(*dims_xyz), axis, degrees, m), "axis_rotation")
    return np.array(m[:], dtype=np.float64).reshape(3, 4)


def affine_invert(m):
    src = (C.c_double * 12)(*np.asarray(m, dtype=np.float64).reshape(-1))
    inv = (C.c_double * 12)()
    _check(lib().orc_affine_invert(src, inv), "affine_invert")
    return np.array(inv[:], dtype=np.float64).reshape(3, 4)


def rotate(vol, axis, degrees):
    vol = _vol(vol)
    out = np.empty_like(vol)
    _check(lib().orc_rotate(_f(vol), _f(out), _dims(vol), axis, degrees), "rotate")
    return out


def attenuate(vol, delta, strict=True):
    vol = _vol(vol)
    out = np.empty_like(vol)
    _check(lib().orc_attenuate(_f(vol), _f(out), _dims(vol), float(delta), int(strict)), "attenuate")
    return out


def sum_image(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return lib().orc_sum(_f(a), a.size)


def norm_image(psf):
    """In place, like Tools.normImage; returns the sum that was divided out."""
    assert psf.dtype == np.float32 and psf.flags.c_contiguous
    return lib().orc_norm_image(_f(psf), psf.size)


def convolve(vol, psf, method="direct", nthreads=0):
    """psf is normalised IN PLACE (reference side effect); returns the convolved volume."""
    vol = _vol(vol)
    assert psf.dtype == np.float32 and psf.flags.c_contiguous and psf.ndim == 3
    out = np.empty_like(vol)
    if method == "direct":
        err = lib().orc_convolve_direct(_f(vol), _dims(vol), _f(psf), _dims(psf), _f(out))
    else:
        nt = nthreads or host_threads()
        err = lib().orc_convolve_fft(_f(vol), _dims(vol), _f(psf), _dims(psf), _f(out), nt)
    _check(err, "convolve")
    return out


def adjust(vol, min_value=0.0001, target_avg=1.0):
    """In place; returns the correction factor."""
    assert vol.dtype == np.float32 and vol.flags.c_contiguous
    return lib().orc_adjust(_f(vol), vol.size, min_value, target_avg)


def extract_slices(vol, inc, snr, seed=464232194):
    vol = _vol(vol)
    z, y, x = vol.shape
    out = np.empty(((z - 1) // inc + 1, y, x), dtype=np.float32)
    _check(lib().orc_extract_slices(_f(vol), _dims(vol), inc, snr, seed, _f(out)), "extract_slices")
    return out


def poisson(a, snr, rnd):
    """In place on a float32 array, drawing from a JavaRandom in flat order."""
    assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().orc_poisson(_f(a), a.size, float(snr), rnd.ptr)


def host_threads():
    """Processors this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which made the
    round-1 reference arm run on one core whenever it was launched under torch.distributed.run."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def simulate_view(gt, psf, axis=0, degrees=15, delta=0.01, min_value=0.0001, target_avg=1.0, inc=3,
                  snr=25.0, seed=464232194, use_fft=True, fft_threads=0, stage_threads=0, poisson_threads=0, want_conv=True):
    """Loop body S/SimulateMultiViewDataset.java:570-585.  Returns (acquired, convolved+adjusted, times).
    fft_threads / stage_threads: 0 = every processor of this process (host_threads()); poisson_threads > 1 is the
    timing-only split of the Poisson loop over slices (see orc_set_poisson_threads)."""
    gt = _vol(gt)
    psf = np.ascontiguousarray(psf, dtype=np.float32).copy()
    z, y, x = gt.shape
    out = np.empty(((z - 1) // inc + 1, y, x), dtype=np.float32)
    conv = np.empty_like(gt) if want_conv else None
    times = (C.c_double * 5)()
    L = lib()
    L.orc_set_stage_threads(stage_threads or host_threads())
    L.orc_set_poisson_threads(poisson_threads)
    try:
        err = L.orc_simulate_view(_f(gt), _dims(gt), _f(psf), _dims(psf), axis, degrees, delta, min_value,
                                  target_avg, inc, snr, seed, int(use_fft), fft_threads or host_threads(),
                                  _f(out), _f(conv) if want_conv else None, times)
    finally:
        L.orc_set_stage_threads(0)
        L.orc_set_poisson_threads(0)
    _check(err, "simulate_view")
    return out, conv, list(times)


def make_isotropic(vol, inc):
    vol = _vol(vol)
    z, y, x = vol.shape
    out = np.empty(((z - 1) * inc + 1, y, x), dtype=np.float32)
    _check(lib().orc_make_isotropic(_f(vol), _dims(vol), inc, _f(out)), "make_isotropic")
    return out


def weight_image(shape_zyx):
    out = np.empty(shape_zyx, dtype=np.float32)
    _check(lib().orc_weight_image(_dims(out), _f(out)), "weight_image")
    return out


def normalize_weights(weights, osem):
    """In place on a list of equal-shape float32 volumes; returns the sum of the normalised weights."""
    n = len(weights)
    fp = C.POINTER(C.c_float)
    arr = (fp * n)(*[_f(w) for w in weights])
    s = np.empty_like(weights[0])
    _check(lib().orc_normalize_weights(arr, n, weights[0].size, osem, _f(s)), "normalize_weights")
    return s


def random_points(n, dims_xyz, seed=535):
    """SimulateBeads.randomPoints with new Random(535) (S/SimulateBeads.java:69,150-166); returns (n, 3) xyz doubles."""
    pts = np.empty((n, 3), dtype=np.float64)
    lib().orc_random_points(n, (C.c_int64 * 3)(*dims_xyz), seed, pts.ctypes.data_as(C.POINTER(C.c_double)))
    return pts


def transform_points(points, dims_xyz, axis, degrees):
    points = np.ascontiguousarray(points, dtype=np.float64)
    out = np.empty_like(points)
    dp = C.POINTER(C.c_double)
    _check(lib().orc_transform_points(points.ctypes.data_as(dp), len(points), (C.c_int64 * 3)(*dims_xyz), axis, degrees,
                                      out.ctypes.data_as(dp)), "transform_points")
    return out


def render_beads(points, sigma_xyz, interval_min_xyz, interval_max_xyz):
    """SimulateBeads.renderPoints for one point list (S/SimulateBeads.java:97-121,168-205).  The image has
    max - min voxels per axis (the reference's own off-by-one at :106); returns it as a (z, y, x) array."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    shape = tuple(int(interval_max_xyz[d] - interval_min_xyz[d]) for d in (2, 1, 0))
    out = np.empty(shape, dtype=np.float32)
    dp = C.POINTER(C.c_double)
    _check(lib().orc_render_beads(points.ctypes.data_as(dp), len(points), (C.c_double * 3)(*sigma_xyz),
                                  (C.c_int64 * 3)(*interval_min_xyz), (C.c_int64 * 3)(*interval_max_xyz), _f(out)), "render_beads")
    return out


def draw_spheres(shape_zyx, scale=2, half_pixel=False, seed=464232194, min_value=0.0, max_value=1.0, max_list=1 << 20):
    """drawSpheres (S/SimulateMultiViewDataset.java:436-522) into a zero volume; returns (volume, list of
    (cx, cy, cz, radius, value) records of the small spheres that were drawn)."""
    img = np.zeros(shape_zyx, dtype=np.float32)
    lst = np.zeros((max_list, 5), dtype=np.float32)
    n = lib().orc_draw_spheres(_f(img), _dims(img), min_value, max_value, scale, int(half_pixel), seed, _f(lst), max_list)
    return img, lst[:min(n, max_list)]


def downsample2x(vol):
    """downSample2x (:394-423)."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    out = np.empty(tuple(s // 2 - 1 for s in vol.shape), dtype=np.float32)
    _check(lib().orc_downsample2x(_f(vol), _dims(vol), _f(out)), "downsample2x")
    return out


def simulate_phantom(size=289, half_pixel=False, seed=464232194):
    """simulate(halfPixelOffset, rnd) (:371-392): size^3 ground truth; returns (volume, number of small spheres)."""
    out = np.empty((size, size, size), dtype=np.float32)
    n = lib().orc_simulate_phantom(size, int(half_pixel), seed, _f(out))
    if n < 0:
        raise RuntimeError("oracle simulate_phantom failed")
    return out, int(n)


def make_square(vol):
    """Tools.makeSquare (S/Tools.java:315-349)."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    m = max(vol.shape)
    out = np.empty((m, m, m), dtype=np.float32)
    _check(lib().orc_make_square(_f(vol), _dims(vol), _f(out)), "make_square")
    return out
