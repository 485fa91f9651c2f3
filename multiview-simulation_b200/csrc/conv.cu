// CUDA backend of the FFT convolution (host logic in fft/conv_driver.h, kernels in fft/line_fft.cuh).
// Replaces convolve() S/SimulateMultiViewDataset.java:253-264 (FFTConvolution on an ExecutorService).
#include <stdlib.h>

#include "ctx.h"
#include "fft/conv_driver.h"
#include "fft/fft_launch.h"

namespace mvsim {

namespace {

struct CudaLauncher {
    mvsim_ctx* ctx;
    bool psf_phase;
    int lanes;

    static int x_blocks(const FftSize& s, int n_rows)
    {
        const int r = x_rows_per_block(s.a, s.b);
        return (n_rows + r - 1) / r;
    }
    int finish(int r, const char* what)
    {
        ctx->launches++;
        if (r == -1) return set_error(ctx, MVSIM_EUNSUPPORTED, "%s: no kernel for this line length", what);
        if (r != 0) return cuda_fail(ctx, (cudaError_t)r, what);
        return MVSIM_OK;
    }
    int launch_x(bool inverse, const FftSize& s, const XParams& q)
    {
        StageTimer t(ctx, psf_phase ? MVSIM_T_PSF : (inverse ? MVSIM_T_FFT_XINV : MVSIM_T_FFT_XFWD));
        return finish(fft_launch(inverse ? FFT_XINV : FFT_XFWD, lanes, s.n, &q, (unsigned)x_blocks(s, q.n_rows), 1, ctx->stream), "x pass");
    }
    int launch_strided(bool inverse, const FftSize& s, const StridedParams& q, int n_outer)
    {
        StageTimer t(ctx, psf_phase ? MVSIM_T_PSF : (inverse ? MVSIM_T_FFT_YINV : MVSIM_T_FFT_YFWD));
        const unsigned tiles = (unsigned)((q.kx_count + lanes - 1) / lanes);
        const unsigned gx = q.swap_grid ? (unsigned)n_outer : tiles, gy = q.swap_grid ? tiles : (unsigned)n_outer;
        return finish(fft_launch(inverse ? FFT_SINV : FFT_SFWD, lanes, s.n, &q, gx, gy, ctx->stream), "strided pass");
    }
    int launch_zfused(const FftSize& s, const ZFusedParams& q, int n_outer)
    {
        StageTimer t(ctx, MVSIM_T_FFT_ZFUSED);
        const unsigned tiles = (unsigned)((q.kx_count + lanes - 1) / lanes);
        return finish(fft_launch(FFT_ZFUSED, lanes, s.n, &q, (unsigned)n_outer, tiles, ctx->stream), "fused z pass");
    }
};

struct Buffers {
    mvsim_ctx* ctx;
    void* p[8];
    int n;
    explicit Buffers(mvsim_ctx* c) : ctx(c), n(0) {}
    ~Buffers() { for (int i = 0; i < n; ++i) dev_free(ctx, p[i]); }
    template <class T> int get(T** out, size_t elems)
    {
        void* q = nullptr;
        int st = dev_alloc(ctx, &q, elems * sizeof(T));
        if (st) return st;
        p[n++] = q;
        *out = static_cast<T*>(q);
        return MVSIM_OK;
    }
};

}  // namespace

int strided_lanes()
{
    static const int lanes = [] {
        const char* e = getenv("MVSIM_LANES");
        return (e && atoi(e) == 4) ? 4 : 8;
    }();
    return lanes;
}

int conv_device(mvsim_ctx* ctx, const float* img, const int64_t dims[3], const float* psf, const int64_t kdims[3],
                float* out, double* d_sum, int keep_inc, int* out_planes)
{
    ConvPlan pl;
    const int perr = make_conv_plan(dims, kdims, &pl);
    if (perr == 1) return set_error(ctx, MVSIM_EINVAL, "convolve: bad dims");
    if (perr) return set_error(ctx, MVSIM_EUNSUPPORTED, "convolve: dim + kdim - 1 exceeds the largest supported FFT line (1600; x: 3200)");
    mvsim_tables tx, ty, tz;
    MVSIM_TRY(get_tables(ctx, pl.sx.n, &tx));
    MVSIM_TRY(get_tables(ctx, pl.sy.n, &ty));
    MVSIM_TRY(get_tables(ctx, pl.sz.n, &tz));

    const int lanes = strided_lanes();
    Buffers buf(ctx);
    ConvWorkspace ws = {};
    MVSIM_TRY(buf.get(&ws.u1, (size_t)pl.u1_elems()));
    MVSIM_TRY(buf.get(&ws.u2, (size_t)pl.u2_elems(lanes)));
    MVSIM_TRY(buf.get(&ws.h, (size_t)pl.h_elems(lanes)));
    MVSIM_TRY(buf.get(&ws.p1, (size_t)pl.p1_elems()));
    MVSIM_TRY(buf.get(&ws.p2, (size_t)pl.p2_elems(lanes)));
    ws.tw_x = tx.tw; ws.twist_x = tx.twist; ws.tw_y = ty.tw; ws.tw_z = tz.tw;

    CudaLauncher l = { ctx, true, lanes };
    MVSIM_TRY(conv_psf_spectrum(l, pl, ws, psf));
    l.psf_phase = false;
    double* partials = nullptr;
    const int planes = conv_out_planes(pl, keep_inc);
    if (out_planes) *out_planes = planes;
    const int nblocks = CudaLauncher::x_blocks(pl.sx, pl.dims[1] * planes);
    if (d_sum) MVSIM_TRY(buf.get(&partials, (size_t)nblocks));
    MVSIM_TRY(conv_apply(l, pl, ws, img, out, partials, keep_inc));
    if (d_sum) {
        StageTimer t(ctx, MVSIM_T_ADJUST);
        MVSIM_TRY(k_sum_partials(ctx, partials, (size_t)nblocks, d_sum));
    }
    return MVSIM_OK;
}

}  // namespace mvsim
