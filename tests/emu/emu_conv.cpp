// CPU emulation of the CUDA line-FFT kernels: runs the very same __host__ __device__ phase
// functions (multiview-simulation_b200/csrc/fft/line_fft.cuh) thread by thread, block by block,
// with a barrier between phases, driven by the same host logic (conv_driver.h).  Built with g++
// by tests/test_emu_conv.py; this is a test of the index arithmetic, not a product path.
#define MVSIM_EMU_SMALL_ONLY 1
#include "fft/conv_driver.h"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace mvsim;

template <class K> static void emulate(const typename K::Params& q, int gx, int gy)
{
    std::vector<float2> sm((size_t)K::SMEM_BYTES / sizeof(float2) + 8);
    std::vector<typename K::State> st(K::THREADS);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            // poison shared memory so stale reads show up
            for (auto& v : sm) { v.x = 1e30f; v.y = -1e30f; }
            for (int t = 0; t < K::THREADS; ++t) K::template phase<0>(q, bx, by, t, sm.data(), st[t]);
            if (K::NPH > 1) for (int t = 0; t < K::THREADS; ++t) K::template phase<1>(q, bx, by, t, sm.data(), st[t]);
            if (K::NPH > 2) for (int t = 0; t < K::THREADS; ++t) K::template phase<2>(q, bx, by, t, sm.data(), st[t]);
            if (K::NPH > 3) for (int t = 0; t < K::THREADS; ++t) K::template phase<3>(q, bx, by, t, sm.data(), st[t]);
            if (K::NPH > 4) for (int t = 0; t < K::THREADS; ++t) K::template phase<4>(q, bx, by, t, sm.data(), st[t]);
            if (K::NPH > 5) for (int t = 0; t < K::THREADS; ++t) K::template phase<5>(q, bx, by, t, sm.data(), st[t]);
        }
}

static constexpr int T = 8;
template <int A, int B> struct Rows { static constexpr int R = (256 / LineShape<A, B>::P) > 0 ? 256 / LineShape<A, B>::P : 1; };

struct EmuLauncher {
    int lanes = T;
    int x_blocks(const FftSize& s, int n_rows) const
    {
        const int p = s.a > s.b ? s.a : s.b;
        const int r = 256 / p > 0 ? 256 / p : 1;
        return (n_rows + r - 1) / r;
    }
    int launch_x(bool inverse, const FftSize& s, const XParams& q)
    {
        const int gx = x_blocks(s, q.n_rows);
        switch (s.n) {
#define MVSIM_X(n_, a_, b_) case n_: if (inverse) emulate<XInv<a_, b_, Rows<a_, b_>::R>>(q, gx, 1); else emulate<XFwd<a_, b_, Rows<a_, b_>::R>>(q, gx, 1); return 0;
            MVSIM_FFT_SIZES(MVSIM_X)
#undef MVSIM_X
        }
        return 5;
    }
    int launch_strided(bool inverse, const FftSize& s, const StridedParams& q, int n_outer)
    {
        const int tiles = (q.kx_count + T - 1) / T;
        const int gx = q.swap_grid ? n_outer : tiles, gy = q.swap_grid ? tiles : n_outer;
        switch (s.n) {
#define MVSIM_X(n_, a_, b_) case n_: if (inverse) emulate<StridedInv<a_, b_, T>>(q, gx, gy); else emulate<StridedFwd<a_, b_, T>>(q, gx, gy); return 0;
            MVSIM_FFT_SIZES(MVSIM_X)
#undef MVSIM_X
        }
        return 5;
    }
    int launch_zfused(const FftSize& s, const ZFusedParams& q, int n_outer)
    {
        const int tiles = (q.kx_count + T - 1) / T;
        switch (s.n) {
#define MVSIM_X(n_, a_, b_) case n_: emulate<ZFused<a_, b_, T>>(q, n_outer, tiles); return 0;
            MVSIM_FFT_SIZES(MVSIM_X)
#undef MVSIM_X
        }
        return 5;
    }
};

extern "C" int emu_plan(const int64_t dims[3], const int64_t kdims[3], int out[9])
{
    ConvPlan pl;
    const int err = make_conv_plan(dims, kdims, &pl);
    if (err) return err;
    out[0] = pl.sx.n; out[1] = pl.sy.n; out[2] = pl.sz.n;
    out[3] = pl.sx.a; out[4] = pl.sx.b; out[5] = pl.sy.a; out[6] = pl.sy.b; out[7] = pl.sz.a; out[8] = pl.sz.b;
    return 0;
}

// psf must already be normalised; sum_out (nullable) receives the sum of the output voxels
// keep_inc > 1: out receives conv_kept_planes(dims, keep_inc) planes (kept slices, then the sum of the others)
extern "C" int emu_convolve(const float* img, const int64_t dims[3], const float* psf, const int64_t kdims[3],
                            float* out, double* sum_out, int keep_inc)
{
    ConvPlan pl;
    int err = make_conv_plan(dims, kdims, &pl);
    if (err) return err;
    std::vector<float2> u1(pl.u1_elems()), u2(pl.u2_elems(T)), h(pl.h_elems(T)), p1(pl.p1_elems()), p2(pl.p2_elems(T));
    std::vector<float2> twx(pl.sx.n), twy(pl.sy.n), twz(pl.sz.n), twist(pl.sx.n);
    fill_twiddles(pl.sx.n, &twx[0].x);
    fill_twiddles(pl.sy.n, &twy[0].x);
    fill_twiddles(pl.sz.n, &twz[0].x);
    fill_twist(pl.sx.n, &twist[0].x);
    // poison the workspaces: every element that is read must have been written by a pass
    for (auto* v : { &u1, &u2, &h, &p1, &p2 })
        for (auto& e : *v) { e.x = 3e30f; e.y = -3e30f; }
    ConvWorkspace ws = { u1.data(), u2.data(), h.data(), p1.data(), p2.data(), twx.data(), twy.data(), twz.data(), twist.data() };
    EmuLauncher l;
    err = conv_psf_spectrum(l, pl, ws, psf);
    if (err) return err;
    const int planes = conv_out_planes(pl, keep_inc);
    std::vector<double> partials(l.x_blocks(pl.sx, pl.dims[1] * planes), 0.0);
    err = conv_apply(l, pl, ws, img, out, partials.data(), keep_inc);
    if (err) return err;
    if (sum_out) { double s = 0; for (double v : partials) s += v; *sum_out = s; }
    return 0;
}
