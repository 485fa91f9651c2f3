"""Shared synthetic inputs for the tests (seeded, small)."""
import numpy as np


def sphere_phantom(shape_zyx, seed=464232194, n_spheres=None):
    """Ball of random small spheres, background 0 -- the statistics of the reference's
    `simulate()` phantom (S/SimulateMultiViewDataset.java:366-522), own generator."""
    z, y, x = shape_zyx
    rng = np.random.default_rng(seed)
    vol = np.zeros(shape_zyx, dtype=np.float32)
    big_r = 0.337 * min(shape_zyx)
    c = np.array([(z - 1) / 2.0, (y - 1) / 2.0, (x - 1) / 2.0])
    n = n_spheres or max(8, int(0.002 * z * y * x / 20))
    for _ in range(n):
        while True:
            p = c + rng.uniform(-big_r, big_r, 3)
            if np.linalg.norm(p - c) <= big_r:
                break
        r = rng.integers(1, 6) / 2.0 + 0.5
        v = np.float32(rng.random())
        lo = np.maximum(np.floor(p - r).astype(int), 0)
        hi = np.minimum(np.ceil(p + r).astype(int) + 1, shape_zyx)
        zz, yy, xx = np.mgrid[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        m = (zz - p[0]) ** 2 + (yy - p[1]) ** 2 + (xx - p[2]) ** 2 <= r * r
        sub = vol[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        sub[m] = np.maximum(sub[m], v)
    return vol


def gaussian_psf(shape_zyx, sigma_zyx, threshold=0.0):
    z, y, x = shape_zyx
    zz, yy, xx = np.mgrid[0:z, 0:y, 0:x].astype(np.float64)
    cz, cy, cx = z // 2, y // 2, x // 2
    g = np.exp(-((zz - cz) ** 2 / (2 * sigma_zyx[0] ** 2) + (yy - cy) ** 2 / (2 * sigma_zyx[1] ** 2)
                 + (xx - cx) ** 2 / (2 * sigma_zyx[2] ** 2)))
    g *= 0.99
    g[g < threshold] = 0
    return np.ascontiguousarray(g, dtype=np.float32)


def rel_err(a, b):
    """max|a-b| / max|b|  -- the tolerance definition of SURVEY.md section 7/8c."""
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))
