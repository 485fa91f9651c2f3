// Sequence of passes of one FFT convolution (host logic, backend independent).
//
// Replaces imglib2-algorithm FFTConvolution.convolve() as called from
// S/SimulateMultiViewDataset.java:257-261 (mirror-single image (x) zero-extended kernel, kernel
// element kdim/2 at the origin, no conjugation, same-size output).  The padded transforms are
// "pruned": each pass reads only rows that carry data and writes only rows the next pass needs.
//
//   PSF:    xfwd(zero ext) -> P1[KZ][KY][KXc] -> y fwd -> P2[KT][KZ][Ny][T] -> z fwd (scaled) -> H[KT][Nz][Ny][T]
//   image:  xfwd(mirror)   -> U1[Z][Y][KXc]  -> y fwd -> U2[KT][Z][Ny][T]  -> z fwd * H, z inv (in place)
//           -> y inv -> U1[Z][Y][KXc] -> x inv + crop (+ per-block sums) -> out[Z][Y][X]
// (KT = kx tiles of T columns; the tile-major layouts keep both strided passes local in memory.)
//
// `L` is the launcher: the CUDA one enqueues kernels on a stream, the emulation one (tests/emu)
// executes the same phase functions on the CPU.
#pragma once
#include "conv_plan.h"
#include "line_fft.cuh"

namespace mvsim {

struct ConvWorkspace {
    float2* u1;     // u1_elems
    float2* u2;     // u2_elems
    float2* h;      // h_elems
    float2* p1;     // p1_elems
    float2* p2;     // p2_elems
    const float2 *tw_x, *tw_y, *tw_z, *twist_x;   // tables for sx.n, sy.n, sz.n
};

// PSF (already normalised to sum 1) -> scaled spectrum ws.h
template <class L> int conv_psf_spectrum(L& l, const ConvPlan& pl, const ConvWorkspace& ws, const float* psf)
{
    const long long kxc = pl.kxc(), T = l.lanes, ny = pl.sy.n;
    XParams xp = {};
    xp.rin = psf; xp.cout = ws.p1; xp.tw = ws.tw_x; xp.twist = ws.twist_x;
    xp.X = pl.kdims[0]; xp.n_rows = pl.kdims[1] * pl.kdims[2]; xp.left = 0; xp.ext = EXT_ZERO;
    int err = l.launch_x(false, pl.sx, xp);
    if (err) return err;

    StridedParams sp = {};      // y: P1 row-major -> P2 tile-major, outer = kz
    sp.in = ws.p1; sp.out = ws.p2; sp.tw = ws.tw_y; sp.kx_count = (int)kxc;
    sp.n_src = pl.kdims[1]; sp.left = 0; sp.ext = EXT_ZERO;
    sp.in_tstride = T; sp.in_estride = kxc; sp.in_ostride = (long long)pl.kdims[1] * kxc;
    sp.out_tstride = (long long)pl.kdims[2] * ny * T; sp.out_estride = T; sp.out_ostride = ny * T;
    sp.swap_grid = 0; sp.scale = 1.0f;
    err = l.launch_strided(false, pl.sy, sp, pl.kdims[2]);
    if (err) return err;

    sp.in = ws.p2; sp.out = ws.h; sp.tw = ws.tw_z;   // z: P2 -> H, both tile-major, outer = ky
    sp.n_src = pl.kdims[2];
    sp.in_tstride = (long long)pl.kdims[2] * ny * T; sp.in_estride = ny * T; sp.in_ostride = T;
    sp.out_tstride = (long long)pl.sz.n * ny * T; sp.out_estride = ny * T; sp.out_ostride = T;
    sp.swap_grid = 1; sp.scale = (float)pl.scale;
    return l.launch_strided(false, pl.sz, sp, pl.sy.n);
}

// Number of z planes conv_apply writes.  keep_inc > 1 (whole-view path, extractSlices keeps every inc-th
// slice, :206): only the kept slices are carried through the inverse y and x passes, plus ONE plane holding
// the sum over all dropped slices -- by linearity of the inverse transforms its voxel sum is exactly what
// adjustImage's mean (S/Tools.java:146) needs from them.  The inverse side shrinks by ~inc.
inline int conv_out_planes(const ConvPlan& pl, int keep_inc)
{
    const int z = pl.dims[2];
    if (keep_inc <= 1) return z;
    const int kept = (z - 1) / keep_inc + 1;
    if (z >= 65536 || keep_inc >= 65536) return z;   // range of the multiply-shift division in the kernel
    return kept + 1 <= z ? kept + 1 : z;       // no room / nothing to gain: plain path
}

// image -> out, using the spectrum in ws.h.  partials (nullable): one double per x-inverse block.
// out has conv_out_planes(pl, keep_inc) planes of Y*X floats.
template <class L> int conv_apply(L& l, const ConvPlan& pl, const ConvWorkspace& ws, const float* img, float* out, double* partials,
                                  int keep_inc = 1)
{
    const int planes = conv_out_planes(pl, keep_inc);
    const bool pruned = planes != pl.dims[2];
    const long long kxc = pl.kxc(), T = l.lanes, ny = pl.sy.n;
    XParams xp = {};
    xp.rin = img; xp.cout = ws.u1; xp.tw = ws.tw_x; xp.twist = ws.twist_x;
    xp.X = pl.dims[0]; xp.n_rows = pl.dims[1] * pl.dims[2]; xp.left = pl.left[0];
    xp.ext = mirror_mode(2 * pl.sx.n, pl.left[0], pl.dims[0]);
    int err = l.launch_x(false, pl.sx, xp);
    if (err) return err;

    StridedParams sp = {};      // y forward: U1 row-major -> U2 tile-major, outer = z
    sp.in = ws.u1; sp.out = ws.u2; sp.tw = ws.tw_y; sp.kx_count = (int)kxc;
    sp.n_src = pl.dims[1]; sp.left = pl.left[1]; sp.ext = mirror_mode(pl.sy.n, pl.left[1], pl.dims[1]);
    sp.in_tstride = T; sp.in_estride = kxc; sp.in_ostride = (long long)pl.dims[1] * kxc;
    sp.out_tstride = (long long)pl.dims[2] * ny * T; sp.out_estride = T; sp.out_ostride = ny * T;
    sp.swap_grid = 0; sp.scale = 1.0f;
    err = l.launch_strided(false, pl.sy, sp, pl.dims[2]);
    if (err) return err;

    ZFusedParams zp = {};
    zp.u = ws.u2; zp.h = ws.h; zp.tw = ws.tw_z; zp.kx_count = (int)kxc;
    zp.n_src = pl.dims[2]; zp.left = pl.left[2]; zp.crop0 = pl.crop0[2];
    zp.ext = mirror_mode(pl.sz.n, pl.left[2], pl.dims[2]);
    zp.keep_inc = pruned ? keep_inc : 1; zp.n_keep = planes - 1; zp.keep_magic = div_magic((uint32_t)zp.keep_inc);
    zp.estride = ny * T; zp.ostride = T;
    zp.u_tstride = (long long)pl.dims[2] * ny * T; zp.h_tstride = (long long)pl.sz.n * ny * T;
    err = l.launch_zfused(pl.sz, zp, pl.sy.n);
    if (err) return err;

    StridedParams ip = {};      // y inverse: U2 tile-major -> U1 row-major
    ip.in = ws.u2; ip.out = ws.u1; ip.tw = ws.tw_y; ip.kx_count = (int)kxc;
    ip.crop0 = pl.crop0[1]; ip.n_out = pl.dims[1];
    ip.in_tstride = (long long)pl.dims[2] * ny * T; ip.in_estride = T; ip.in_ostride = ny * T;
    ip.out_tstride = T; ip.out_estride = kxc; ip.out_ostride = (long long)pl.dims[1] * kxc;
    ip.swap_grid = 0; ip.scale = 1.0f;
    err = l.launch_strided(true, pl.sy, ip, planes);
    if (err) return err;

    XParams ix = {};
    ix.cin = ws.u1; ix.rout = out; ix.tw = ws.tw_x; ix.twist = ws.twist_x; ix.partials = partials;
    ix.X = pl.dims[0]; ix.n_rows = pl.dims[1] * planes; ix.crop0 = pl.crop0[0];
    return l.launch_x(true, pl.sx, ix);
}

}  // namespace mvsim
