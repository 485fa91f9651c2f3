/*
 * mvsim_oracle.c -- CPU restatement of the per-view acquisition pipeline of
 * PreibischLab/multiview-simulation.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker the CUDA path is compared against.  It is never
 * linked into, imported by or called from the product library (libmvsim.so);
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.
 *
 * PARITY UNPINNED: the reference ships no tests, no golden outputs and cannot
 * be built here (no JVM, and the arithmetic lives in un-vendored jars:
 * imglib2, imglib2-algorithm 0.18.3 fft2.FFTConvolution, mpicbg 1.6.6
 * AffineModel3D / RealSum, Mines JTK).  Every function below restates the
 * semantics visible at the reference's own call sites (cited as
 * S/<file>:<line>, S = src/main/java/net/preibisch/simulation) plus the
 * published behaviour of those libraries; it is pinned only by the
 * known-answer tests in tests/test_oracle_*.py that follow from that source.
 *
 * Layout everywhere: ImgLib2 ArrayImg order, x fastest: idx = x + X*(y + Y*z).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads used by the non-FFT stages (the reference runs them on one thread,
 * S/SimulateMultiViewDataset.java:104-135,318-364); 0 = all cores (checker use). */
static int g_stage_threads = 0;
void orc_set_stage_threads(int n) { g_stage_threads = n; }
static int stage_threads(void)
{
#ifdef _OPENMP
    return g_stage_threads > 0 ? g_stage_threads : omp_get_max_threads();
#else
    return 1;
#endif
}
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* processors available to this process, independent of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) */
int orc_hw_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
/* Timing-only knob of bench.py's "best effort" baseline: with n > 1 extractSlices gives every kept slice its own
 * java.util.Random (seed + slice index) so that the O(lambda) loops of different slices run on n threads.  The
 * reference draws all slices from ONE sequential generator (S/SimulateMultiViewDataset.java:76,183): 0 / 1 (the
 * default, and what every test uses) reproduces that order. */
static int g_poisson_threads = 0;
void orc_set_poisson_threads(int n) { g_poisson_threads = n; }

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_ENOMEM 2

/* ------------------------------------------------------------------------ */
/* java.util.Random (JDK specification; used at S/SimulateMultiViewDataset.java:76,
 * S/uncommons/PoissonGenerator.java:101).  48-bit LCG, bit-exact.            */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t s; } orc_jrandom;

void orc_jrandom_init(orc_jrandom* r, int64_t seed)
{
    r->s = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
}

static inline int32_t jnext(orc_jrandom* r, int bits)
{
    r->s = (r->s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(int64_t)(r->s >> (48 - bits));
}

int32_t orc_jrandom_next_int(orc_jrandom* r) { return jnext(r, 32); }

int64_t orc_jrandom_next_long(orc_jrandom* r)
{
    /* ((long)next(32) << 32) + next(32), both signed */
    int64_t hi = (int64_t)jnext(r, 32);
    int64_t lo = (int64_t)jnext(r, 32);
    return (int64_t)((uint64_t)hi << 32) + lo;
}

double orc_jrandom_next_double(orc_jrandom* r)
{
    int64_t a = (int64_t)jnext(r, 26);
    int64_t b = (int64_t)jnext(r, 27);
    return (double)((a << 27) + b) * 0x1.0p-53;
}

/* Random.nextInt(int bound), JDK specification (power-of-two shortcut, rejection loop otherwise) */
int32_t orc_jrandom_next_int_bound(orc_jrandom* r, int32_t bound)
{
    int32_t v = jnext(r, 31);
    const int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)v) >> 31);
    for (int32_t u = v; (int32_t)((uint32_t)u - (uint32_t)(v = u % bound) + (uint32_t)m) < 0; u = jnext(r, 31)) {}
    return v;
}

/* ------------------------------------------------------------------------ */
/* a1: axisRotation  S/SimulateMultiViewDataset.java:80-102                   */
/* 3x4 row-major doubles m00..m23 (mpicbg AffineModel3D).                      */
/* ------------------------------------------------------------------------ */
static void affine_preconcat(double* t, const double* a)
{
    /* t <- a * t   (mpicbg preConcatenate) */
    double r[12];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            r[i * 4 + j] = a[i * 4 + 0] * t[0 * 4 + j] + a[i * 4 + 1] * t[1 * 4 + j] + a[i * 4 + 2] * t[2 * 4 + j];
        r[i * 4 + 3] = a[i * 4 + 0] * t[3] + a[i * 4 + 1] * t[7] + a[i * 4 + 2] * t[11] + a[i * 4 + 3];
    }
    memcpy(t, r, sizeof(r));
}

int orc_axis_rotation(const int64_t dims[3], int axis, int degrees, double m[12])
{
    if (axis < 0 || axis > 2) return ORC_EINVAL;
    /* ( in.max(d) - in.min(d) ) / 2 with long division: :84-86 */
    double c[3];
    for (int d = 0; d < 3; ++d) c[d] = (double)((dims[d] - 1) / 2);
    double t1[12] = { 1, 0, 0, -c[0], 0, 1, 0, -c[1], 0, 0, 1, -c[2] };
    double t2[12] = { 1, 0, 0, c[0], 0, 1, 0, c[1], 0, 0, 1, c[2] };
    /* (float)Math.toRadians(degrees): :90 ; toRadians = deg / 180.0 * PI */
    const double rad_d = (double)degrees / 180.0 * 3.14159265358979323846;
    const double th = (double)(float)rad_d;
    const double cs = cos(th), sn = sin(th);
    double rot[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
    if (axis == 0) { rot[5] = cs; rot[6] = -sn; rot[9] = sn; rot[10] = cs; }
    else if (axis == 1) { rot[0] = cs; rot[2] = sn; rot[8] = -sn; rot[10] = cs; }
    else { rot[0] = cs; rot[1] = -sn; rot[4] = sn; rot[5] = cs; }
    memcpy(m, t1, sizeof(t1));
    affine_preconcat(m, rot);   /* :98 */
    affine_preconcat(m, t2);    /* :99 */
    return ORC_OK;
}

/* mpicbg AffineModel3D.createInverse(): cofactor inverse, t' = -M^-1 t */
int orc_affine_invert(const double m[12], double inv[12])
{
    const double m00 = m[0], m01 = m[1], m02 = m[2], m03 = m[3];
    const double m10 = m[4], m11 = m[5], m12 = m[6], m13 = m[7];
    const double m20 = m[8], m21 = m[9], m22 = m[10], m23 = m[11];
    const double det = m00 * m11 * m22 + m10 * m21 * m02 + m20 * m01 * m12
                     - m02 * m11 * m20 - m12 * m21 * m00 - m22 * m01 * m10;
    if (det == 0) return ORC_EINVAL;
    const double idet = 1.0 / det;
    inv[0] = (m11 * m22 - m12 * m21) * idet;
    inv[1] = (m02 * m21 - m01 * m22) * idet;
    inv[2] = (m01 * m12 - m02 * m11) * idet;
    inv[4] = (m12 * m20 - m10 * m22) * idet;
    inv[5] = (m00 * m22 - m02 * m20) * idet;
    inv[6] = (m02 * m10 - m00 * m12) * idet;
    inv[8] = (m10 * m21 - m11 * m20) * idet;
    inv[9] = (m01 * m20 - m00 * m21) * idet;
    inv[10] = (m00 * m11 - m01 * m10) * idet;
    inv[3] = -inv[0] * m03 - inv[1] * m13 - inv[2] * m23;
    inv[7] = -inv[4] * m03 - inv[5] * m13 - inv[6] * m23;
    inv[11] = -inv[8] * m03 - inv[9] * m13 - inv[10] * m23;
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* a2: rotateAroundAxis  S/SimulateMultiViewDataset.java:104-135              */
/* imglib2 NLinearInterpolator3D over Views.extendZero: floor + double       */
/* weights, each tap rounded to float (FloatType.mul(double)), float adds in  */
/* the order 000,100,110,010,011,111,101,001 (x,y,z bits).                    */
/* ------------------------------------------------------------------------ */
static inline float tap0(const float* in, const int64_t d[3], int64_t x, int64_t y, int64_t z)
{
    if (x < 0 || y < 0 || z < 0 || x >= d[0] || y >= d[1] || z >= d[2]) return 0.0f;
    return in[x + d[0] * (y + d[1] * z)];
}

static inline int64_t mirror_single(int64_t i, int64_t n)
{
    if (n == 1) return 0;
    const int64_t p = 2 * (n - 1);
    int64_t j = i % p;
    if (j < 0) j += p;
    return j >= n ? p - j : j;
}

static inline float tapm(const float* in, const int64_t d[3], int64_t x, int64_t y, int64_t z)
{
    return in[mirror_single(x, d[0]) + d[0] * (mirror_single(y, d[1]) + d[1] * mirror_single(z, d[2]))];
}

/* mode 0: zero extension, mode 1: mirror-single extension */
static inline float nlinear3(const float* in, const int64_t d[3], const double p[3], int mode)
{
    const double f0 = floor(p[0]), f1 = floor(p[1]), f2 = floor(p[2]);
    const int64_t x = (int64_t)f0, y = (int64_t)f1, z = (int64_t)f2;
    const double w0 = p[0] - f0, w0i = 1.0 - w0;
    const double w1 = p[1] - f1, w1i = 1.0 - w1;
    const double w2 = p[2] - f2, w2i = 1.0 - w2;
    float (*tap)(const float*, const int64_t*, int64_t, int64_t, int64_t) = mode ? tapm : tap0;
    float acc = (float)((double)tap(in, d, x, y, z) * (w0i * w1i * w2i));
    acc += (float)((double)tap(in, d, x + 1, y, z) * (w0 * w1i * w2i));
    acc += (float)((double)tap(in, d, x + 1, y + 1, z) * (w0 * w1 * w2i));
    acc += (float)((double)tap(in, d, x, y + 1, z) * (w0i * w1 * w2i));
    acc += (float)((double)tap(in, d, x, y + 1, z + 1) * (w0i * w1 * w2));
    acc += (float)((double)tap(in, d, x + 1, y + 1, z + 1) * (w0 * w1 * w2));
    acc += (float)((double)tap(in, d, x + 1, y, z + 1) * (w0 * w1i * w2));
    acc += (float)((double)tap(in, d, x, y, z + 1) * (w0i * w1i * w2));
    return acc;
}

int orc_rotate(const float* in, float* out, const int64_t dims[3], int axis, int degrees)
{
    double fwd[12], inv[12];
    if (orc_axis_rotation(dims, axis, degrees, fwd)) return ORC_EINVAL;
    if (orc_affine_invert(fwd, inv)) return ORC_EINVAL;
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
#pragma omp parallel for collapse(2) schedule(static) num_threads(stage_threads())
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t x = 0; x < X; ++x) {
                /* mpicbg applyInPlace: l0*m00 + l1*m01 + l2*m02 + m03 */
                const double l0 = (double)x, l1 = (double)y, l2 = (double)z;
                double p[3];
                p[0] = l0 * inv[0] + l1 * inv[1] + l2 * inv[2] + inv[3];
                p[1] = l0 * inv[4] + l1 * inv[5] + l2 * inv[6] + inv[7];
                p[2] = l0 * inv[8] + l1 * inv[9] + l2 * inv[10] + inv[11];
                out[x + X * (y + Y * z)] = nlinear3(in, dims, p, 0);
            }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* a3: attenuate3d  S/SimulateMultiViewDataset.java:318-364                   */
/* strict != 0 reproduces the reference's loop bound dimension(0) (:345);     */
/* it is only defined for X <= Y (X > Y walks out of bounds in the reference).*/
/* ------------------------------------------------------------------------ */
int orc_attenuate(const float* in, float* out, const int64_t dims[3], double delta, int strict)
{
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    const int64_t steps = strict ? X : Y;
    if (steps > Y) return ORC_EINVAL;
    memset(out, 0, sizeof(float) * (size_t)(X * Y * Z));
#pragma omp parallel for collapse(2) schedule(static) num_threads(stage_threads())
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t x = 0; x < X; ++x) {
            double n = 1.0;
            int64_t y = Y - 1;
            for (int64_t s = 0; s < steps; ++s, --y) {
                const double v = (double)in[x + X * (y + Y * z)];
                const double phi = v * delta * n;       /* :350 */
                n = fmax(n - phi, 0.0);                 /* :353 */
                out[x + X * (y + Y * z)] = (float)(v * n);  /* :356 */
            }
        }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* a4: Tools.sumImage / normImage  S/Tools.java:112-132                       */
/* mpicbg RealSum = accurate double summation; restated as Neumaier           */
/* compensated summation in flat (x fastest) order.                           */
/* ------------------------------------------------------------------------ */
double orc_sum(const float* img, size_t n)
{
    double s = 0.0, c = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double v = (double)img[i];
        const double t = s + v;
        if (fabs(s) >= fabs(v)) c += (s - t) + v; else c += (v - t) + s;
        s = t;
    }
    return s + c;
}

double orc_norm_image(float* psf, size_t n)
{
    const double sum = orc_sum(psf, n);
    for (size_t i = 0; i < n; ++i) psf[i] = (float)((double)psf[i] / sum);
    return sum;
}

/* ------------------------------------------------------------------------ */
/* a5: convolve  S/SimulateMultiViewDataset.java:253-264                      */
/* Definition (independent of any padded FFT size):                           */
/*   out[x] = sum_k psfN[k] * I_ms[x - (k - c)],  c_d = kdim_d / 2,           */
/* I_ms = mirror-single extension, psfN = psf normalised in place (:255).     */
/* ------------------------------------------------------------------------ */
int orc_convolve_direct(const float* img, const int64_t dims[3], float* psf, const int64_t kdims[3], float* out)
{
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    const int64_t KX = kdims[0], KY = kdims[1], KZ = kdims[2];
    if (X < 1 || Y < 1 || Z < 1 || KX < 1 || KY < 1 || KZ < 1) return ORC_EINVAL;
    orc_norm_image(psf, (size_t)(KX * KY * KZ));
    const int64_t cx = KX / 2, cy = KY / 2, cz = KZ / 2;
    int64_t* mx = (int64_t*)malloc(sizeof(int64_t) * (size_t)((X + KX) + (Y + KY) + (Z + KZ)));
    if (!mx) return ORC_ENOMEM;
    int64_t* my = mx + (X + KX);
    int64_t* mz = my + (Y + KY);
    /* table index t = (x - (k - c)) + (K - 1 - c)  in [0, n + K - 1) */
    for (int64_t t = 0; t < X + KX - 1; ++t) mx[t] = mirror_single(t - (KX - 1 - cx), X);
    for (int64_t t = 0; t < Y + KY - 1; ++t) my[t] = mirror_single(t - (KY - 1 - cy), Y);
    for (int64_t t = 0; t < Z + KZ - 1; ++t) mz[t] = mirror_single(t - (KZ - 1 - cz), Z);
#pragma omp parallel for collapse(2) schedule(static) num_threads(stage_threads())
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t x = 0; x < X; ++x) {
                double acc = 0.0;
                for (int64_t kz = 0; kz < KZ; ++kz) {
                    const int64_t sz = mz[z - kz + KZ - 1];
                    for (int64_t ky = 0; ky < KY; ++ky) {
                        const int64_t sy = my[y - ky + KY - 1];
                        const float* row = img + X * (sy + Y * sz);
                        const float* krow = psf + KX * (ky + KY * kz);
                        const int64_t* mrow = mx + x + KX - 1;
                        for (int64_t kx = 0; kx < KX; ++kx)
                            acc += (double)krow[kx] * (double)row[mrow[-kx]];
                    }
                }
                out[x + X * (y + Y * z)] = (float)acc;
            }
    free(mx);
    return ORC_OK;
}

/* ---- float32 FFT path, shaped like imglib2-algorithm fft2.FFTConvolution ----
 * (S/SimulateMultiViewDataset.java:257-261): image extended mirror-single and
 * centred in a padded interval of "fast" size >= dim + kdim - 1, kernel
 * zero-extended with element kdim/2 at the origin (periodic wrap), float32
 * transforms, spectrum product without conjugation, inverse scaled by 1/N,
 * original interval copied out.  The transforms here are complex-to-complex
 * mixed radix (2,3,5,7 + generic) in float32 with double-built twiddles; the
 * 1-D lines are distributed over `nthreads` threads like the reference's
 * ExecutorService.  Result agrees with orc_convolve_direct to float32 FFT
 * round-off (checked in tests/test_oracle_conv.py).                          */

typedef struct { float re, im; } cpx;

static int64_t next_smooth(int64_t n)
{
    for (;; ++n) {
        int64_t m = n;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        while (m % 7 == 0) m /= 7;
        if (m == 1) return n;
    }
}

int64_t orc_fft_size(int64_t n) { return next_smooth(n < 1 ? 1 : n); }

/* recursive decimation-in-time mixed radix; tw = table of exp(-+2 pi i k / N0) for the
 * top-level N0, stride walks the table. */
static void fft_rec(const cpx* in, cpx* out, int64_t n, int64_t istride, const cpx* tw, int64_t twstride, cpx* scratch)
{
    if (n == 1) { out[0] = in[0]; return; }
    int64_t r = 0;
    if (n % 4 == 0) r = 4; else if (n % 2 == 0) r = 2; else if (n % 3 == 0) r = 3;
    else if (n % 5 == 0) r = 5; else if (n % 7 == 0) r = 7; else r = n;
    const int64_t m = n / r;
    for (int64_t q = 0; q < r; ++q)
        fft_rec(in + q * istride, out + q * m, m, istride * r, tw, twstride * r, scratch);
    /* combine: X[k + m*j] = sum_q W_n^{q(k + m j)} Y_q[k] */
    for (int64_t k = 0; k < m; ++k) {
        cpx t[16];
        cpx* tt = r <= 16 ? t : scratch;
        for (int64_t q = 0; q < r; ++q) {
            const cpx w = tw[(q * k * twstride)];
            const cpx y = out[q * m + k];
            tt[q].re = y.re * w.re - y.im * w.im;
            tt[q].im = y.re * w.im + y.im * w.re;
        }
        for (int64_t j = 0; j < r; ++j) {
            float sr = 0.f, si = 0.f;
            for (int64_t q = 0; q < r; ++q) {
                /* W_r^{q j} = tw[(q*j % r) * m * twstride] */
                const cpx w = tw[((q * j) % r) * m * twstride];
                sr += tt[q].re * w.re - tt[q].im * w.im;
                si += tt[q].re * w.im + tt[q].im * w.re;
            }
            scratch[r + j].re = sr; scratch[r + j].im = si;
        }
        for (int64_t j = 0; j < r; ++j) out[k + m * j] = scratch[r + j];
    }
}

typedef struct { int64_t n; cpx* tw; } fft_plan;

static int plan_init(fft_plan* p, int64_t n, int sign)
{
    p->n = n;
    p->tw = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    if (!p->tw) return ORC_ENOMEM;
    for (int64_t k = 0; k < n; ++k) {
        const double a = (double)sign * 2.0 * 3.14159265358979323846 * (double)k / (double)n;
        p->tw[k].re = (float)cos(a); p->tw[k].im = (float)sin(a);
    }
    return ORC_OK;
}

/* transform all lines along `axis` of a complex volume of dims n[3] (x fastest) */
static int fft_axis(cpx* vol, const int64_t n[3], int axis, int sign, int nthreads)
{
    fft_plan p;
    if (plan_init(&p, n[axis], sign)) return ORC_ENOMEM;
    const int64_t len = n[axis];
    const int64_t stride = axis == 0 ? 1 : (axis == 1 ? n[0] : n[0] * n[1]);
    const int64_t a = axis == 0 ? n[1] : n[0];
    const int64_t b = axis == 2 ? n[1] : n[2];
    const int64_t sa = axis == 0 ? n[0] : 1;
    const int64_t sb = axis == 2 ? n[0] : n[0] * n[1];
    int err = 0;
    (void)nthreads;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        cpx* lin = (cpx*)malloc(sizeof(cpx) * (size_t)len * 4 + sizeof(cpx) * 64);
        if (!lin) {
#pragma omp atomic write
            err = 1;
        } else {
            cpx* lout = lin + len;
            cpx* scratch = lout + len;
#pragma omp for collapse(2) schedule(static)
            for (int64_t j = 0; j < b; ++j)
                for (int64_t i = 0; i < a; ++i) {
                    cpx* base = vol + i * sa + j * sb;
                    for (int64_t t = 0; t < len; ++t) lin[t] = base[t * stride];
                    fft_rec(lin, lout, len, 1, p.tw, 1, scratch);
                    for (int64_t t = 0; t < len; ++t) base[t * stride] = lout[t];
                }
            free(lin);
        }
    }
    free(p.tw);
    return err ? ORC_ENOMEM : ORC_OK;
}

int orc_convolve_fft(const float* img, const int64_t dims[3], float* psf, const int64_t kdims[3], float* out, int nthreads)
{
    int64_t n[3], off[3];
    for (int d = 0; d < 3; ++d) {
        if (dims[d] < 1 || kdims[d] < 1) return ORC_EINVAL;
        n[d] = next_smooth(dims[d] + kdims[d] - 1);
        off[d] = (n[d] - dims[d]) / 2;          /* image centred in the padded interval */
    }
    orc_norm_image(psf, (size_t)(kdims[0] * kdims[1] * kdims[2]));
    const size_t total = (size_t)(n[0] * n[1] * n[2]);
    cpx* a = (cpx*)malloc(sizeof(cpx) * total);
    cpx* k = (cpx*)calloc(total, sizeof(cpx));
    if (!a || !k) { free(a); free(k); return ORC_ENOMEM; }
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t z = 0; z < n[2]; ++z)
        for (int64_t y = 0; y < n[1]; ++y) {
            const int64_t sz = mirror_single(z - off[2], dims[2]);
            const int64_t sy = mirror_single(y - off[1], dims[1]);
            for (int64_t x = 0; x < n[0]; ++x) {
                const int64_t sx = mirror_single(x - off[0], dims[0]);
                cpx v = { img[sx + dims[0] * (sy + dims[1] * sz)], 0.f };
                a[x + n[0] * (y + n[1] * z)] = v;
            }
        }
    for (int64_t z = 0; z < kdims[2]; ++z)
        for (int64_t y = 0; y < kdims[1]; ++y)
            for (int64_t x = 0; x < kdims[0]; ++x) {
                const int64_t px = ((x - kdims[0] / 2) % n[0] + n[0]) % n[0];
                const int64_t py = ((y - kdims[1] / 2) % n[1] + n[1]) % n[1];
                const int64_t pz = ((z - kdims[2] / 2) % n[2] + n[2]) % n[2];
                k[px + n[0] * (py + n[1] * pz)].re = psf[x + kdims[0] * (y + kdims[1] * z)];
            }
    int err = 0;
    for (int d = 0; d < 3 && !err; ++d) err = fft_axis(a, n, d, -1, nthreads);
    for (int d = 0; d < 3 && !err; ++d) err = fft_axis(k, n, d, -1, nthreads);
    if (!err) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
        for (size_t i = 0; i < total; ++i) {
            const cpx u = a[i], v = k[i];
            a[i].re = u.re * v.re - u.im * v.im;
            a[i].im = u.re * v.im + u.im * v.re;
        }
        for (int d = 0; d < 3 && !err; ++d) err = fft_axis(a, n, d, +1, nthreads);
    }
    if (!err) {
        const float scale = (float)(1.0 / (double)total);
        for (int64_t z = 0; z < dims[2]; ++z)
            for (int64_t y = 0; y < dims[1]; ++y)
                for (int64_t x = 0; x < dims[0]; ++x)
                    out[x + dims[0] * (y + dims[1] * z)] =
                        a[(x + off[0]) + n[0] * ((y + off[1]) + n[1] * (z + off[2]))].re * scale;
    }
    free(a); free(k);
    return err;
}

/* ------------------------------------------------------------------------ */
/* a6: Tools.adjustImage  S/Tools.java:143-159                                */
/* ------------------------------------------------------------------------ */
double orc_adjust(float* img, size_t n, float min_value, float target_avg)
{
    const double avg = orc_sum(img, n) / (double)n;
    /* ( targetAverage - minValue ) is a float subtraction, widened for the divide */
    const double corr = (double)(float)(target_avg - min_value) / avg;
    for (size_t i = 0; i < n; ++i) img[i] = (float)((double)img[i] * corr);
    for (size_t i = 0; i < n; ++i) img[i] = img[i] + min_value;
    return corr;
}

/* ------------------------------------------------------------------------ */
/* a8: Tools.poissonProcess  S/Tools.java:73-86 + PoissonGenerator.nextValue  */
/* S/uncommons/PoissonGenerator.java:95-109; java.util.Random draw order.     */
/* Deviation: lambda < 0 or NaN never terminates in the reference; */
/* the oracle returns 0 there.                                                */
/* ------------------------------------------------------------------------ */
void orc_poisson(float* img, size_t n, double snr, orc_jrandom* rnd)
{
    const double q = snr / sqrt(5.0);
    const double mul = pow(q, 2.0);
    for (size_t i = 0; i < n; ++i) {
        const double lambda = (double)img[i] * mul;
        int x = 0;
        if (lambda > 0.0) {
            double t = 0.0;
            for (;;) {
                t -= log(orc_jrandom_next_double(rnd)) / lambda;
                if (t > 1.0) break;
                ++x;
            }
        } else if (lambda == 0.0) {
            /* one draw is consumed: t -= log(u)/0 -> +inf (or NaN -> never; u==0,lambda==0) */
            (void)orc_jrandom_next_double(rnd);
        }
        img[i] = (float)x;
    }
}

/* ------------------------------------------------------------------------ */
/* a7: extractSlices  S/SimulateMultiViewDataset.java:195-231                 */
/* out dims (X, Y, (Z-1)/inc + 1); snr < 0 -> raw copy (:211-214).            */
/* ------------------------------------------------------------------------ */
int orc_extract_slices(const float* in, const int64_t dims[3], int inc, float snr, int64_t seed, float* out)
{
    if (inc < 1) return ORC_EINVAL;
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    if (g_poisson_threads > 1 && snr >= 0.0f) {
        const int64_t nk = (Z - 1) / inc + 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_poisson_threads)
        for (int64_t k = 0; k < nk; ++k) {
            orc_jrandom r;
            orc_jrandom_init(&r, seed + k);
            float* o = out + X * Y * k;
            memcpy(o, in + X * Y * (k * inc), sizeof(float) * (size_t)(X * Y));
            orc_poisson(o, (size_t)(X * Y), (double)snr, &r);
        }
        return ORC_OK;
    }
    orc_jrandom rnd;
    orc_jrandom_init(&rnd, seed);
    int64_t cz = 0;
    for (int64_t z = 0; z < Z; z += inc, ++cz) {
        float* o = out + X * Y * cz;
        memcpy(o, in + X * Y * z, sizeof(float) * (size_t)(X * Y));
        if (snr >= 0.0f) orc_poisson(o, (size_t)(X * Y), (double)snr, &rnd);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* "next" row f-1: post-acquisition chain of main()                            */
/* makeIsotropic  S/SimulateMultiViewDataset.java:144-171: linear z up-sampling */
/* by inc over Views.extendMirrorSingle, position (x, y, (float)z/(float)inc).  */
/* ------------------------------------------------------------------------ */
int orc_make_isotropic(const float* in, const int64_t dims[3], int inc, float* out)
{
    if (inc < 1) return ORC_EINVAL;
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    const int64_t ZO = (Z - 1) * inc + 1;
#pragma omp parallel for collapse(2) schedule(static) num_threads(stage_threads())
    for (int64_t z = 0; z < ZO; ++z)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t x = 0; x < X; ++x) {
                double p[3] = { (double)x, (double)y, (double)((float)z / (float)inc) };   /* :163-165 */
                out[x + X * (y + Y * z)] = nlinear3(in, dims, p, 1);
            }
    return ORC_OK;
}

/* computeWeightImage  S/SimulateMultiViewDataset.java:280-316 (cosine taper over 40 px along y; delta unused) */
int orc_weight_image(const int64_t dims[3], float* out)
{
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    const int cosine_span = 40;
    const int size_y = (int)Y;
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < Y; ++y) {
            const int l = size_y - (int)y - 1;
            float v;
            if (l < size_y / 2) v = 1.0f;
            else if (l > size_y / 2 + cosine_span) v = 0.0f;
            else {
                const double pos = ((double)(l - size_y / 2) / (double)cosine_span) * 3.14159265358979323846;
                v = (float)((cos(pos) + 1.0) / 2.0);
            }
            for (int64_t x = 0; x < X; ++x) out[x + X * (y + Y * z)] = v;
        }
    return ORC_OK;
}

/* weight normalisation of main()  S/SimulateMultiViewDataset.java:615-661: in place over n_views volumes,
 * w_v = min(1, osem * (w_v / sum)); sum_out (nullable) = sum of the normalised weights */
int orc_normalize_weights(float** w, int n_views, size_t n, float osem, float* sum_out)
{
    if (n_views < 1) return ORC_EINVAL;
    for (size_t i = 0; i < n; ++i) {
        float sum = 0;
        for (int v = 0; v < n_views; ++v) sum += w[v][i];
        float s2 = 0;
        for (int v = 0; v < n_views; ++v) {
            if (sum == 0) w[v][i] = 0;
            else w[v][i] = fminf(1.0f, osem * (w[v][i] / (float)sum));
            s2 += w[v][i];
        }
        if (sum_out) sum_out[i] = s2;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* "next" row f-2: bead phantom renderer  S/SimulateBeads.java:97-205          */
/* randomPoints (:150-166), transformPoints = axisRotation applied to points   */
/* (:131-148), renderPoints/addGaussian (:97-121,168-205): every point inside  */
/* the interval adds (float)prod_d exp(-x_d^2 / 2 sigma_d^2) * 1000f over a box */
/* of 2*getSuggestedKernelDiameter(sigma_d) voxels per axis, in point order.    */
/* ------------------------------------------------------------------------ */
static int suggested_kernel_diameter(double sigma)
{
    /* imglib2 Util.getSuggestedKernelDiameter: max(3, 2*(int)(3 sigma + 0.5) + 1) */
    int size = 3;
    if (sigma > 0) size = 2 * (int)(3.0 * sigma + 0.5) + 1;
    return size < 3 ? 3 : size;
}

void orc_random_points(int n, const int64_t range_dims[3], int64_t seed, double* points /* n*3 */)
{
    orc_jrandom r;
    orc_jrandom_init(&r, seed);
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)     /* range.min = 0, range.max = dim - 1  (FinalInterval(dims)) */
            points[3 * i + d] = orc_jrandom_next_double(&r) * (double)(range_dims[d] - 1) + 0.0;
}

int orc_transform_points(const double* points, int n, const int64_t range_dims[3], int axis, int degrees, double* out)
{
    double m[12];
    if (orc_axis_rotation(range_dims, axis, degrees, m)) return ORC_EINVAL;
    for (int i = 0; i < n; ++i) {
        const double* p = points + 3 * i;
        for (int r = 0; r < 3; ++r) out[3 * i + r] = p[0] * m[4 * r] + p[1] * m[4 * r + 1] + p[2] * m[4 * r + 2] + m[4 * r + 3];
    }
    return ORC_OK;
}

/* renderPoints (:97-121): image dims = interval.max - interval.min (ONE LESS than the interval's dimension, as written
 * at :106), a point is kept when 0 <= p - min <= interval.dimension - 1 = max - min (isInsideAdjust :123-135, which also
 * shifts the point by -min).  out has prod(max - min) floats. */
int orc_render_beads(const double* points, int n, const double sigma[3], const int64_t imin[3], const int64_t imax[3], float* out)
{
    int64_t dims[3];
    for (int d = 0; d < 3; ++d) { dims[d] = imax[d] - imin[d]; if (dims[d] < 1) return ORC_EINVAL; }
    const int64_t X = dims[0], Y = dims[1], Z = dims[2];
    memset(out, 0, sizeof(float) * (size_t)(X * Y * Z));
    int size[3];
    double two_sq[3];
    for (int d = 0; d < 3; ++d) { size[d] = suggested_kernel_diameter(sigma[d]) * 2; two_sq[d] = 2 * sigma[d] * sigma[d]; }
    for (int i = 0; i < n; ++i) {
        double p[3];
        int inside = 1;
        for (int d = 0; d < 3; ++d) {
            p[d] = points[3 * i + d] - (double)imin[d];
            if (p[d] < 0 || p[d] > (double)(imax[d] - imin[d] + 1 - 1)) { inside = 0; break; }
        }
        if (!inside) continue;
        int64_t mn[3];
        for (int d = 0; d < 3; ++d) mn[d] = (int64_t)(int)floor(p[d] + 0.5) - size[d] / 2;     /* (int)Math.round(location) */
        for (int64_t z = mn[2]; z < mn[2] + size[2]; ++z)
            for (int64_t y = mn[1]; y < mn[1] + size[1]; ++y)
                for (int64_t x = mn[0]; x < mn[0] + size[0]; ++x) {
                    if (x < 0 || y < 0 || z < 0 || x >= X || y >= Y || z >= Z) continue;      /* Views.extendZero: writes outside are lost */
                    double value = 1;
                    const double dx = p[0] - (double)x, dy = p[1] - (double)y, dz = p[2] - (double)z;
                    value *= exp(-(dx * dx) / two_sq[0]);
                    value *= exp(-(dy * dy) / two_sq[1]);
                    value *= exp(-(dz * dz) / two_sq[2]);
                    float* o = out + x + X * (y + Y * z);
                    *o = *o + ((float)value * 1000.0f);
                }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* "next" row f-3: sphere phantom  S/SimulateMultiViewDataset.java:366-522     */
/* imglib2-algorithm HyperSphereCursor (third party, not under the reference  */
/* tree; restated from the published source): z, then y, then x ascending;    */
/* NESTED integer radii  ry = (long)sqrt(R^2 - dz^2), rx = (long)sqrt(ry^2 - dy^2). */
/* ------------------------------------------------------------------------ */
typedef void (*sphere_visit)(int64_t x, int64_t y, int64_t z, void* user);

static void hypersphere_for_each(const int64_t c[3], int64_t R, sphere_visit fn, void* user)
{
    for (int64_t dz = -R; dz <= R; ++dz) {
        const int64_t ry = (int64_t)sqrt((double)(R * R - dz * dz));
        for (int64_t dy = -ry; dy <= ry; ++dy) {
            const int64_t rx = (int64_t)sqrt((double)(ry * ry - dy * dy));
            for (int64_t dx = -rx; dx <= rx; ++dx) fn(c[0] + dx, c[1] + dy, c[2] + dz, user);
        }
    }
}

typedef struct { float* img; int64_t dims[3]; float v; } paint_ctx;

static void paint_max(int64_t x, int64_t y, int64_t z, void* user)
{
    paint_ctx* c = (paint_ctx*)user;
    if (x < 0 || y < 0 || z < 0 || x >= c->dims[0] || y >= c->dims[1] || z >= c->dims[2]) return;   /* cannot happen for the reference sizes */
    float* o = c->img + x + c->dims[0] * (y + c->dims[1] * z);
    /* value.setReal(Math.max(randomValue, value.getRealDouble())) (:517): float store of a double max == max of the float roundings */
    if (c->v > *o) *o = c->v;
}

typedef struct { orc_jrandom* rnd; paint_ctx paint; int scale; int half_pixel; double minv, maxv; int64_t mod; int64_t n_small; float* list; int64_t list_cap; } draw_ctx;

static void draw_visit(int64_t x, int64_t y, int64_t z, void* user)
{
    draw_ctx* d = (draw_ctx*)user;
    const int radius = orc_jrandom_next_int_bound(d->rnd, 10 * d->scale) + 1;                 /* :485 */
    double rv = orc_jrandom_next_double(d->rnd);                                              /* :506 */
    const int64_t rounded = (int64_t)floor(rv * 10000 + 0.5);                                 /* Math.round */
    if (rounded % d->mod != 0) return;                                                        /* :509 */
    rv = orc_jrandom_next_double(d->rnd) * (d->maxv - d->minv) + d->minv;                     /* :512 */
    int64_t c[3] = { x, y, z };
    if (d->half_pixel) { c[0] += 1; c[1] += 1; }                                              /* :496-499: all but the last dimension */
    if (d->list && d->n_small < d->list_cap) {
        float* e = d->list + 5 * d->n_small;
        e[0] = (float)c[0]; e[1] = (float)c[1]; e[2] = (float)c[2]; e[3] = (float)radius; e[4] = (float)rv;
    }
    d->n_small++;
    d->paint.v = (float)rv;
    hypersphere_for_each(c, radius, paint_max, &d->paint);
}

/* drawSpheres (:436-522) into a zeroed img of dims; returns the number of small spheres drawn.  list (nullable) receives
 * up to list_cap records (cx, cy, cz, radius, value). */
int64_t orc_draw_spheres(float* img, const int64_t dims[3], double minv, double maxv, int scale, int half_pixel, int64_t seed,
                         float* list, int64_t list_cap)
{
    orc_jrandom r;
    orc_jrandom_init(&r, seed);
    int64_t c[3], min_size = dims[0];
    for (int d = 0; d < 3; ++d) { c[d] = dims[d] / 2; if (dims[d] < min_size) min_size = dims[d]; }
    const int64_t R = min_size / 2 - 47 * scale - 1;                                          /* :462 */
    draw_ctx d;
    d.rnd = &r; d.paint.img = img; memcpy(d.paint.dims, dims, sizeof(d.paint.dims)); d.scale = scale; d.half_pixel = half_pixel;
    d.minv = minv; d.maxv = maxv; d.mod = (int64_t)(7 * scale) * (7 * scale) * (7 * scale);  /* Util.pow(7*scale, 3) */
    d.n_small = 0; d.list = list; d.list_cap = list_cap;
    if (R < 0) return 0;
    hypersphere_for_each(c, R, draw_visit, &d);
    return d.n_small;
}

/* downSample2x (:394-423): dims/2 - 1, n-linear at 2l + 0.5 over the mirror-single extension */
int orc_downsample2x(const float* in, const int64_t dims[3], float* out)
{
    int64_t od[3];
    for (int d = 0; d < 3; ++d) { od[d] = dims[d] / 2 - 1; if (od[d] < 1) return ORC_EINVAL; }
    const int nt = stage_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t z = 0; z < od[2]; ++z)
        for (int64_t y = 0; y < od[1]; ++y)
            for (int64_t x = 0; x < od[0]; ++x) {
                const double p[3] = { (double)x * 2.0 + 0.5, (double)y * 2.0 + 0.5, (double)z * 2.0 + 0.5 };
                out[x + od[0] * (y + od[1] * z)] = nlinear3(in, dims, p, 1);
            }
    return ORC_OK;
}

/* simulate(halfPixelOffset, rnd) (:371-392) for a given final size (reference: 289); out has size^3 floats */
int64_t orc_simulate_phantom(int size, int half_pixel, int64_t seed, float* out)
{
    const int scale = 2;
    const int64_t big = (int64_t)(size + 1) * scale;
    const int64_t dims[3] = { big, big, big };
    float* img = (float*)calloc((size_t)(big * big * big), sizeof(float));
    if (!img) return -1;
    const int64_t n = orc_draw_spheres(img, dims, 0.0, 1.0, scale, half_pixel, seed, NULL, 0);
    const int st = orc_downsample2x(img, dims, out);
    free(img);
    return st == ORC_OK ? n : -1;
}

/* Tools.makeSquare (S/Tools.java:315-349): cube of the largest dimension, input centred with the offset
 * -square/2 + dim/2 (integer divisions), padded with the input's minimum */
int orc_make_square(const float* in, const int64_t dims[3], float* out)
{
    int64_t m = 0;
    for (int d = 0; d < 3; ++d) if (dims[d] > m) m = dims[d];
    float mn = FLT_MAX;
    const size_t n = (size_t)(dims[0] * dims[1] * dims[2]);
    for (size_t i = 0; i < n; ++i) mn = in[i] < mn ? in[i] : mn;
    for (int64_t z = 0; z < m; ++z)
        for (int64_t y = 0; y < m; ++y)
            for (int64_t x = 0; x < m; ++x) {
                const int64_t sx = x - m / 2 + dims[0] / 2, sy = y - m / 2 + dims[1] / 2, sz = z - m / 2 + dims[2] / 2;
                const int inside = sx >= 0 && sy >= 0 && sz >= 0 && sx < dims[0] && sy < dims[1] && sz < dims[2];
                out[x + m * (y + m * z)] = inside ? in[sx + dims[0] * (sy + dims[1] * sz)] : mn;
            }
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* The per-view loop body S/SimulateMultiViewDataset.java:570-585, used by     */
/* smoke() and bench.py's cpu_baseline leg.  times[5] (seconds): rotate,       */
/* attenuate, convolve, adjust, extract+poisson.                              */
/* ------------------------------------------------------------------------ */
static double now_s(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

int orc_simulate_view(const float* gt, const int64_t dims[3], float* psf, const int64_t kdims[3],
                      int axis, int degrees, double delta, float min_value, float target_avg,
                      int inc, float snr, int64_t seed, int use_fft, int nthreads,
                      float* out, float* conv_out, double* times)
{
    const size_t n = (size_t)(dims[0] * dims[1] * dims[2]);
    float* a = (float*)malloc(sizeof(float) * n);
    float* b = (float*)malloc(sizeof(float) * n);
    if (!a || !b) { free(a); free(b); return ORC_ENOMEM; }
    int err;
    double t0 = now_s();
    err = orc_rotate(gt, a, dims, axis, degrees);
    double t1 = now_s();
    if (!err) err = orc_attenuate(a, b, dims, delta, 1);
    double t2 = now_s();
    if (!err) err = use_fft ? orc_convolve_fft(b, dims, psf, kdims, a, nthreads)
                            : orc_convolve_direct(b, dims, psf, kdims, a);
    double t3 = now_s();
    if (!err) orc_adjust(a, n, min_value, target_avg);
    double t4 = now_s();
    if (!err && conv_out) memcpy(conv_out, a, sizeof(float) * n);
    if (!err) err = orc_extract_slices(a, dims, inc, snr, seed, out);
    double t5 = now_s();
    if (times) { times[0] = t1 - t0; times[1] = t2 - t1; times[2] = t3 - t2; times[3] = t4 - t3; times[4] = t5 - t4; }
    free(a); free(b);
    return err;
}
