// Sequence of passes of one FFT convolution (host logic, backend independent).
//
// Replaces imglib2-algorithm FFTConvolution.convolve() as called from
// S/SimulateMultiViewDataset.java:257-261 (mirror-single image (x) zero-extended kernel, kernel
// element kdim/2 at the origin, no conjugation, same-size output).  The padded transforms are
// "pruned": each pass reads only rows that carry data and writes only rows the next pass needs.
//
//   PSF:    xfwd(zero ext) -> P1[KZ][KY][KXc] -> y fwd -> P2[tiles][KZ][Ny][T] -> z fwd (scaled) -> H[tiles][Nz][Ny][T]
//   image:  xfwd(mirror)   -> U1[Zl][Y][KXc]
//           per y block:  y fwd -> U2[KT][Zl][Ny][T]  ==(all-to-all when world > 1)==>  [S][tiles][Zl][Ny][T]
//                         z fwd * H, z inv (in place)  ==(all-to-all back)==>  U2 -> y inv -> U1'[planes][Y][KXc]
//           x inv + crop (+ per-block sums) -> out[planes][Y][X]
// KT = kx tiles of T columns.  On one GPU (world = 1) Zl = Z, tiles = KT, S = 1 and there is no exchange.
// Slab decomposition (largest single volume): a rank owns Zl = Z/world planes for the x and y passes and
// KT/world tiles (with all planes) for the z pass; U2 viewed as [world][tiles][Zl][Ny][T] is at once the
// all-to-all SEND layout (destination major, contiguous equal chunks) and, after the exchange, the
// segmented layout the fused z pass addresses directly -- pack and unpack are folded into the passes.
//
// `L` is the launcher: the CUDA one enqueues kernels on a stream, the emulation one (tests/emu)
// executes the same phase functions on the CPU.
#pragma once
#include "conv_plan.h"
#include "line_fft.cuh"

namespace mvsim {

struct ConvWorkspace {
    float2* u1;     // u1_elems(z_local): forward x spectra
    float2* u1o;    // inverse-side rows; may alias u1 when there is a single y block
    float2* u2;     // u2_elems(T, z_local): y-pass output = exchange send buffer
    float2* ex;     // exchange receive buffer (== u2 when world == 1)
    float2* h;      // h_elems(T, tiles_own)
    float2* p1;     // p1_elems
    float2* p2;     // p2_elems(T, tiles_own)
    const float2 *tw_x, *tw_y, *tw_z, *twist_x;   // tables for sx.n, sy.n, sz.n
    // peer-to-peer slab mode (n_peers = world > 1): z-pass input buffers [S][tiles][Zl][Ny][T] and inverse-side buffers
    // [KT][Zl][Ny][T] of EVERY rank (own entries are local pointers); the y forward pass and the fused z pass store
    // straight into the owners' buffers, so ex == peers_x[rank] and u2 == peers_y[rank]
    int n_peers;
    float2* peers_x[kMaxRanks];
    float2* peers_y[kMaxRanks];
};

// PSF (already normalised to sum 1) -> scaled spectrum ws.h for this rank's tiles
template <class L> int conv_psf_spectrum(L& l, const ConvPlan& pl, const SlabGeom& g, const ConvWorkspace& ws, const float* psf)
{
    const long long kxc = pl.kxc(), T = l.lanes, ny = pl.sy.n;
    XParams xp = {};
    xp.rin = psf; xp.cout = ws.p1; xp.tw = ws.tw_x; xp.twist = ws.twist_x;
    xp.X = pl.kdims[0]; xp.n_rows = pl.kdims[1] * pl.kdims[2]; xp.left = 0; xp.ext = EXT_ZERO; xp.esz = (unsigned)sizeof(float);
    int err = l.launch_x(false, pl.sx, xp);
    if (err) return err;

    StridedParams sp = {};      // y: P1 row-major -> P2 tile-major (own tiles), outer = kz
    sp.in = ws.p1; sp.out = ws.p2; sp.tw = ws.tw_y; sp.kx_count = (int)kxc;
    sp.n_src = pl.kdims[1]; sp.left = 0; sp.ext = EXT_ZERO;
    sp.in_tstride = T; sp.in_estride = kxc; sp.in_ostride = (long long)pl.kdims[1] * kxc;
    sp.out_tstride = (long long)pl.kdims[2] * ny * T; sp.out_estride = T; sp.out_ostride = ny * T;
    sp.swap_grid = 0; sp.scale = l.h_on_the_fly ? (float)pl.scale : 1.0f;
    sp.tile0 = g.tile0; sp.in_tile_global = 1; sp.out_tile_global = 0;
    err = l.launch_strided(false, pl.sy, sp, g.tiles_own, pl.kdims[2]);
    if (err) return err;
    if (l.h_on_the_fly) return 0;       // the fused z pass transforms P2 along z itself: no H is written

    sp.in = ws.p2; sp.out = ws.h; sp.tw = ws.tw_z;   // z: P2 -> H, both tile-major, outer = ky
    sp.n_src = pl.kdims[2];
    sp.in_tstride = (long long)pl.kdims[2] * ny * T; sp.in_estride = ny * T; sp.in_ostride = T;
    sp.out_tstride = (long long)pl.sz.n * ny * T; sp.out_estride = ny * T; sp.out_ostride = T;
    sp.swap_grid = 1; sp.scale = (float)pl.scale;
    sp.in_tile_global = 0; sp.out_tile_global = 0;
    return l.launch_strided(false, pl.sz, sp, g.tiles_own, pl.sy.n);
}

// x forward of this rank's planes: img [Zl][Y][X] -> ws.u1
template <class L> int conv_forward_x(L& l, const ConvPlan& pl, const SlabGeom& g, const ConvWorkspace& ws, const float* img)
{
    XParams xp = {};
    xp.rin = img; xp.cout = ws.u1; xp.tw = ws.tw_x; xp.twist = ws.twist_x;
    xp.X = pl.dims[0]; xp.n_rows = pl.dims[1] * g.z_local; xp.left = pl.left[0];
    xp.ext = mirror_mode(2 * pl.sx.n, pl.left[0], pl.dims[0]); xp.esz = (unsigned)sizeof(float);
    return l.launch_x(false, pl.sx, xp);
}

// y forward of block b: ws.u1 -> ws.u2 [KT][Zl][Ny][T]
template <class L> int conv_forward_y(L& l, const ConvPlan& pl, const SlabGeom& g, const ConvWorkspace& ws, int b)
{
    const long long kxc = pl.kxc(), T = l.lanes, ny = pl.sy.n;
    StridedParams sp = {};
    sp.in = ws.u1; sp.out = ws.u2; sp.tw = ws.tw_y; sp.kx_count = (int)kxc;
    sp.n_src = pl.dims[1];
    sp.left = pl.left[1] - b * pl.y_block;       // padded index q of block b holds source row q - left[1] + b*y_block
    // the mirror may be evaluated branch free when every padded index folds at most once
    {
        const int lo = -sp.left, hi = pl.sy.n - 1 - sp.left;
        sp.ext = (pl.dims[1] > 1 && lo >= -(pl.dims[1] - 1) && hi <= 2 * (pl.dims[1] - 1)) ? EXT_MIRROR1 : EXT_MIRROR_GENERAL;
    }
    sp.in_tstride = T; sp.in_estride = kxc; sp.in_ostride = (long long)pl.dims[1] * kxc;
    sp.out_tstride = (long long)g.z_local * ny * T; sp.out_estride = T; sp.out_ostride = ny * T;
    sp.swap_grid = 0; sp.scale = 1.0f;
    sp.tile0 = 0; sp.in_tile_global = 1; sp.out_tile_global = 1;
    if (ws.n_peers > 1) {
        sp.n_peers = ws.n_peers; sp.my_rank = g.rank; sp.peer_tiles = g.tiles_own; sp.peer_tiles_magic = div_magic((uint32_t)g.tiles_own);
        for (int r = 0; r < ws.n_peers; ++r) sp.out_peers[r] = ws.peers_x[r];
    }
    return l.launch_strided(false, pl.sy, sp, g.tiles_total, g.z_local);
}

// Number of z planes the inverse side carries.  keep_inc > 1 (whole-view path on one GPU, extractSlices keeps
// every inc-th slice, :206): only the kept slices go through the inverse y and x passes, plus ONE plane
// holding the sum over all dropped slices -- by linearity of the inverse transforms its voxel sum is exactly
// what adjustImage's mean (S/Tools.java:146) needs from them.  The inverse side shrinks by ~inc.
inline int conv_out_planes(const ConvPlan& pl, const SlabGeom& g, int keep_inc)
{
    const int z = pl.dims[2];
    if (g.world > 1) return g.z_local;
    if (keep_inc <= 1) return z;
    const int kept = (z - 1) / keep_inc + 1;
    if (z >= 65536 || keep_inc >= 65536) return z;   // range of the multiply-shift division in the kernel
    return kept + 1 <= z ? kept + 1 : z;       // no room / nothing to gain: plain path
}

// fused z pass on this rank's tiles, in place on buf = [S][tiles][Zl][Ny][T]
// sum_is_total (nullable): set to true when the extra plane of the pruned inverse side carries the sum of ALL cropped planes (the
// polyphase kernel) instead of the sum of the dropped ones -- the inverse x pass must then count that plane alone
template <class L> int conv_middle_z(L& l, const ConvPlan& pl, const SlabGeom& g, const ConvWorkspace& ws, float2* buf, int keep_inc,
                                     bool* sum_is_total = nullptr)
{
    if (sum_is_total) *sum_is_total = false;
    const long long T = l.lanes, ny = pl.sy.n;
    const int planes = conv_out_planes(pl, g, keep_inc);
    const bool pruned = g.world == 1 && planes != pl.dims[2];
    ZFusedParams zp = {};
    zp.u = buf; zp.h = ws.h; zp.tw = ws.tw_z; zp.kx_count = (int)pl.kxc();
    zp.n_src = pl.dims[2]; zp.left = pl.left[2]; zp.crop0 = pl.crop0[2];
    zp.ext = mirror_mode(pl.sz.n, pl.left[2], pl.dims[2]);
    zp.tile0 = g.tile0; zp.zg = g.z_local; zp.zg_magic = div_magic((uint32_t)g.z_local);
    zp.keep_inc = pruned ? keep_inc : 1; zp.n_keep = planes - 1; zp.keep_magic = div_magic((uint32_t)zp.keep_inc);
    zp.estride = ny * T; zp.ostride = T;
    zp.estride32 = (g.world == 1 && ws.n_peers <= 1 && (long long)pl.sz.n * ny * T < 0x0fffffffLL) ? (int)(ny * T) : 0;
    zp.u_tstride = (long long)g.z_local * ny * T; zp.seg_stride = (long long)g.tiles_own * g.z_local * ny * T;
    zp.h_tstride = (long long)pl.sz.n * ny * T;
    if (pl.dims[2] >= 65536) return 5;
    if (l.h_on_the_fly) {
        zp.h_mode = 1; zp.p2 = ws.p2; zp.k_src = pl.kdims[2]; zp.p2_tstride = (long long)pl.kdims[2] * ny * T;
    }
    if (ws.n_peers > 1) {
        zp.n_peers = ws.n_peers;
        for (int r = 0; r < ws.n_peers; ++r) zp.out_peers[r] = ws.peers_y[r];
    }
    // polyphase form (ZFusedPoly): inc phases of sz.n / inc points, sum plane from the time domain.  Needs a mirror extension that
    // reaches every padded index and an image line at least as long as the PSF line (the launcher may veto: line length, A/B knob)
    if (pruned && l.h_on_the_fly && zp.estride32 != 0 && zp.ext == EXT_MIRROR1 && zfused_poly_ok(pl.sz.n, keep_inc) && pl.dims[2] >= pl.kdims[2] &&
        l.z_polyphase(pl.sz, keep_inc, pl.kdims[2])) {
        const int perr = l.launch_zfused_poly(pl.sz, zp, g.tiles_own, pl.sy.n, keep_inc);
        if (perr != -2) {                   // -2: the launcher could not set the kernel up (no TMA descriptor): spectral kernels below
            if (sum_is_total) *sum_is_total = true;
            return perr;
        }
    }
    // decimated inverse when the kept planes are whole columns of the exchange (the launcher may veto: line length, A/B knob)
    if (pruned && l.h_on_the_fly && zp.estride32 != 0 && zfused_dec_ok(pl.sz.a, pl.sz.b, keep_inc) && l.z_decimate(pl.sz))
        return l.launch_zfused_dec(pl.sz, zp, g.tiles_own, pl.sy.n, keep_inc);
    return l.launch_zfused(pl.sz, zp, g.tiles_own, pl.sy.n);
}

// D'[k] = sum over the cropped outputs n in [crop0, crop0 + n_src) that are NOT kept (kept: n = crop0 + inc m) of exp(+2 pi i n k / N),
// k < N.  ZFusedDec: the sum of the dropped slices of the unscaled inverse transform is sum_k Yhat[k] D'[k].  Exact integer phase
// reduction (n k mod N) into a double sine / cosine table; stored as float2.
inline void zfused_dec_table(int n, int crop0, int n_src, int inc, float2* out)
{
    const double two_pi = 6.283185307179586476925286766559;
    double* cs = new double[2 * (size_t)n];
    for (int m = 0; m < n; ++m) { cs[2 * m] = cos(two_pi * (double)m / (double)n); cs[2 * m + 1] = sin(two_pi * (double)m / (double)n); }
    for (int k = 0; k < n; ++k) {
        double re = 0.0, im = 0.0;
        for (int o = 0; o < n_src; ++o) {
            if (inc > 1 && o % inc == 0) continue;
            const long long ph = ((long long)(crop0 + o) * k) % n;
            re += cs[2 * ph]; im += cs[2 * ph + 1];
        }
        out[k].x = (float)re; out[k].y = (float)im;
    }
    delete[] cs;
}

// y inverse of block b: ws.u2 (tile-major [KT][Zl][Ny][T], `planes` of the Zl planes in use) -> rows of ws.u1o
template <class L> int conv_inverse_y(L& l, const ConvPlan& pl, const SlabGeom& g, const ConvWorkspace& ws, int b, int planes)
{
    const long long kxc = pl.kxc(), T = l.lanes, ny = pl.sy.n;
    const int y0 = b * pl.y_block;
    StridedParams ip = {};
    ip.in = ws.u2; ip.out = ws.u1o; ip.tw = ws.tw_y; ip.kx_count = (int)kxc;
    ip.crop0 = pl.crop0[1];
    ip.n_out = pl.dims[1] - y0 < pl.y_block ? pl.dims[1] - y0 : pl.y_block;
    ip.out_offset = y0;
    ip.in_tstride = (long long)g.z_local * ny * T; ip.in_estride = T; ip.in_ostride = ny * T;
    // peer-to-peer slab mode: the fused z pass delivered [KT][Ny][Zl][T] (each owner's z range of a line in one bulk copy)
    if (ws.n_peers > 1) { ip.in_estride = (long long)g.z_local * T; ip.in_ostride = T; }
    ip.out_tstride = T; ip.out_estride = kxc; ip.out_ostride = (long long)pl.dims[1] * kxc;
    ip.swap_grid = 0; ip.scale = 1.0f;
    ip.tile0 = 0; ip.in_tile_global = 1; ip.out_tile_global = 1;
    return l.launch_strided(true, pl.sy, ip, g.tiles_total, planes);
}

// x inverse + crop: ws.u1o -> out [planes][Y][X]; partials (nullable): one double per block
// sum_last_plane_only: the per-block sums count the LAST plane alone (it carries the sum of all cropped planes, see conv_middle_z)
template <class L> int conv_inverse_x(L& l, const ConvPlan& pl, const ConvWorkspace& ws, float* out, double* partials, int planes,
                                      bool sum_last_plane_only = false)
{
    XParams ix = {};
    ix.cin = ws.u1o; ix.rout = out; ix.tw = ws.tw_x; ix.twist = ws.twist_x; ix.partials = partials;
    ix.sum_row0 = sum_last_plane_only ? (long long)pl.dims[1] * (planes - 1) : 0;
    ix.X = pl.dims[0]; ix.n_rows = pl.dims[1] * planes; ix.crop0 = pl.crop0[0]; ix.esz = (unsigned)sizeof(float);
    return l.launch_x(true, pl.sx, ix);
}

// Whole convolution on one GPU (world == 1): image -> out, using the spectrum in ws.h.
// out has conv_out_planes(pl, g, keep_inc) planes of Y*X floats.
template <class L> int conv_apply(L& l, const ConvPlan& pl, const SlabGeom& g, const ConvWorkspace& ws, const float* img, float* out,
                                  double* partials, int keep_inc = 1)
{
    const int planes = conv_out_planes(pl, g, keep_inc);
    int err = conv_forward_x(l, pl, g, ws, img);
    bool sum_total = false;         // the same kernel serves every y block, so the last answer stands for all
    for (int b = 0; b < pl.y_blocks && !err; ++b) {
        err = conv_forward_y(l, pl, g, ws, b);
        if (!err) err = conv_middle_z(l, pl, g, ws, ws.u2, keep_inc, &sum_total);
        if (!err) err = conv_inverse_y(l, pl, g, ws, b, planes);
    }
    if (!err) err = conv_inverse_x(l, pl, ws, out, partials, planes, sum_total);
    return err;
}

}  // namespace mvsim
