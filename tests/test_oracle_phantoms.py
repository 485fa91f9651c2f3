"""CPU oracle of the input generators: known answers derived from the cited source lines
and the JDK specification of java.util.Random.  No GPU."""
import math

import numpy as np


def test_java_random_next_int_bound_known_values(oracle):
    # widely published JDK sequences
    r = oracle.JavaRandom(42)
    assert [r.next_int_bound(10) for _ in range(10)] == [0, 3, 8, 4, 0, 5, 5, 8, 9, 3]
    r = oracle.JavaRandom(0)
    assert [r.next_int_bound(100) for _ in range(5)] == [60, 48, 29, 47, 15]
    r = oracle.JavaRandom(7)
    assert all(0 <= r.next_int_bound(16) < 16 for _ in range(100))          # power-of-two branch


def test_random_points_follow_next_double_in_xyz_order(oracle):
    pts = oracle.random_points(5, (512, 512, 200), seed=535)
    r = oracle.JavaRandom(535)
    for i in range(5):
        for d, n in enumerate((512, 512, 200)):
            assert pts[i, d] == r.next_double() * (n - 1)                     # range.max - range.min = dim - 1 (:161)


def test_transform_points_is_axis_rotation(oracle):
    pts = oracle.random_points(20, (64, 48, 40))
    m = oracle.axis_rotation((64, 48, 40), 0, 45)
    out = oracle.transform_points(pts, (64, 48, 40), 0, 45)
    ref = pts @ m[:, :3].T + m[:, 3]
    assert np.allclose(out, ref, rtol=0, atol=1e-12)
    assert np.array_equal(out[:, 0], pts[:, 0])                               # axis 0 leaves x alone
    assert np.allclose(oracle.transform_points(pts, (64, 48, 40), 0, 0), pts, atol=1e-12)


def test_single_bead_is_the_analytic_gaussian_times_1000(oracle):
    sigma = (1.0, 1.0, 3.0)
    p = np.array([[20.3, 17.8, 25.5]])
    img = oracle.render_beads(p, sigma, (0, 0, 0), (48, 40, 56))
    assert img.shape == (56, 40, 48)                                          # max - min per axis (:106), z y x
    # support: 2 * getSuggestedKernelDiameter(sigma) = 2 * max(3, 2*(int)(3 sigma + .5) + 1) = 14, 14, 38 voxels
    nz = np.argwhere(img > 0)
    lo, hi = nz.min(0), nz.max(0)
    assert tuple(hi - lo + 1) == (38, 14, 14)
    assert tuple(lo) == (26 - 19, 18 - 7, 20 - 7)                             # round(p) - size/2
    z, y, x = 24, 18, 21
    v = math.exp(-(20.3 - x) ** 2 / 2) * math.exp(-(17.8 - y) ** 2 / 2) * math.exp(-(25.5 - z) ** 2 / 18)
    assert abs(img[z, y, x] - np.float32(np.float32(v) * np.float32(1000))) <= 1e-4


def test_beads_outside_interval_are_dropped_and_interval_min_shifts(oracle):
    sigma = (1.0, 1.0, 1.0)
    pts = np.array([[-0.5, 5, 5], [5, 5, 5], [31.5, 5, 5], [5, 5, 40.0]])
    img = oracle.render_beads(pts, sigma, (0, 0, 0), (31, 31, 31))
    only = oracle.render_beads(pts[1:2], sigma, (0, 0, 0), (31, 31, 31))
    assert np.array_equal(img, only)
    # a point at exactly max (= dimension - 1) is still inside although the image is one voxel shorter (:106 vs :129)
    edge = oracle.render_beads(np.array([[31.0, 5, 5]]), sigma, (0, 0, 0), (31, 31, 31))
    assert edge.max() > 0 and edge.shape == (31, 31, 31)
    shifted = oracle.render_beads(pts[1:2] + 10.0, sigma, (10, 10, 10), (41, 41, 41))
    assert np.array_equal(shifted, only)


def test_overlapping_beads_add_in_point_order(oracle):
    sigma = (1.0, 1.0, 1.0)
    a, b = np.array([[10.2, 10.0, 10.0]]), np.array([[11.1, 10.4, 9.7]])
    both = oracle.render_beads(np.vstack([a, b]), sigma, (0, 0, 0), (24, 24, 24))
    ia, ib = oracle.render_beads(a, sigma, (0, 0, 0), (24, 24, 24)), oracle.render_beads(b, sigma, (0, 0, 0), (24, 24, 24))
    assert np.array_equal(both, ia + ib)                                      # float32 adds, first a then b


def test_hypersphere_order_and_membership(oracle):
    """drawSpheres on a volume just large enough for a big sphere of radius 2 (scale 1: R = min/2 - 47 - 1)."""
    dims = 100
    img, lst = oracle.draw_spheres((dims, dims, dims), scale=1, seed=1)
    # R = 50 - 48 = 2: nested radii give 1 + 9 + 13 + 9 + 1 voxels; P(select) ~ 1/343*... so usually nothing is drawn
    assert img.shape == (dims, dims, dims)
    r = oracle.JavaRandom(1)
    n_sel = 0
    for _ in range(33):
        r.next_int_bound(10)
        v = r.next_double()
        if math.floor(v * 10000 + 0.5) % 343 == 0:
            r.next_double()
            n_sel += 1
    assert len(lst) == n_sel


def test_sphere_phantom_statistics_and_replay(oracle):
    vol, n = oracle.simulate_phantom(size=121, seed=464232194)
    assert vol.shape == (121, 121, 121) and n > 0
    assert vol.min() == 0.0 and 0.0 < vol.max() < 1.0
    # big sphere radius at 2x: 244/2 - 95 = 27 -> 13.5 px after down-sampling, small spheres reach 10 px further
    zz, yy, xx = np.mgrid[0:121, 0:121, 0:121]
    rad = np.sqrt((zz - 60.5) ** 2 + (yy - 60.5) ** 2 + (xx - 60.5) ** 2)
    assert np.all(vol[rad > 13.5 + 10 + 1.5] == 0)
    again, n2 = oracle.simulate_phantom(size=121, seed=464232194)
    assert n2 == n and np.array_equal(vol, again)
    other, _ = oracle.simulate_phantom(size=121, seed=5)
    assert not np.array_equal(vol, other)


def test_small_spheres_follow_the_nested_radius_rule(oracle):
    big, lst = oracle.draw_spheres((242, 242, 242), scale=2, seed=464232194)
    assert len(lst) > 3
    # repaint from the list with an independent numpy implementation of HyperSphereCursor membership
    ref = np.zeros_like(big)
    for cx, cy, cz, rad, val in lst:
        cx, cy, cz, rad = int(cx), int(cy), int(cz), int(rad)
        for dz in range(-rad, rad + 1):
            ry = math.isqrt(rad * rad - dz * dz)
            for dy in range(-ry, ry + 1):
                rx = math.isqrt(ry * ry - dy * dy)
                row = ref[cz + dz, cy + dy, cx - rx:cx + rx + 1]
                np.maximum(row, np.float32(val), out=row)
    assert np.array_equal(big, ref)
    assert all(1 <= r <= 20 for r in lst[:, 3])


def test_half_pixel_offset_shifts_x_and_y_by_one(oracle):
    a, la = oracle.draw_spheres((242, 242, 242), scale=2, seed=9, half_pixel=False)
    b, lb = oracle.draw_spheres((242, 242, 242), scale=2, seed=9, half_pixel=True)
    assert len(la) == len(lb) > 0
    assert np.array_equal(lb[:, :2], la[:, :2] + 1) and np.array_equal(lb[:, 2:], la[:, 2:])
    assert np.array_equal(b[:, 1:, 1:], a[:, :-1, :-1])


def test_downsample2x_is_the_mean_of_the_2x2x2_block_at_odd_offsets(oracle):
    v = np.random.default_rng(3).random((12, 14, 16), dtype=np.float32)
    out = oracle.downsample2x(v)
    assert out.shape == (5, 6, 7)
    blk = sum(v[dz:10 + dz:2, dy:12 + dy:2, dx:14 + dx:2].astype(np.float64) for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)) / 8
    assert np.allclose(out, blk, rtol=0, atol=2e-7)
    const = oracle.downsample2x(np.full((8, 8, 8), 0.75, dtype=np.float32))
    assert np.all(const == 0.75)


def test_make_square_pads_with_minimum_and_centres(oracle):
    v = np.random.default_rng(4).random((5, 9, 6), dtype=np.float32) + 0.5
    v[2, 3, 1] = 0.125
    sq = oracle.make_square(v)
    assert sq.shape == (9, 9, 9)
    oz, oy, ox = 9 // 2 - 5 // 2, 0, 9 // 2 - 6 // 2
    assert np.array_equal(sq[oz:oz + 5, oy:oy + 9, ox:ox + 6], v)
    mask = np.ones_like(sq, dtype=bool)
    mask[oz:oz + 5, oy:oy + 9, ox:ox + 6] = False
    assert np.all(sq[mask] == np.float32(0.125))
    cube = np.random.default_rng(5).random((4, 4, 4), dtype=np.float32)
    assert np.array_equal(oracle.make_square(cube), cube)
