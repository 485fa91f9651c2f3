// Two-level (N = A*B) line FFTs held in registers, one shared-memory exchange per direction.
//
// Replaces the 1-D float32 FFT lines that imglib2-algorithm fft2.FFTConvolution hands to Mines JTK
// (reference call site S/SimulateMultiViewDataset.java:257-261; S = src/main/java/net/preibisch/simulation).
//
// Decomposition (n = n1*B + n2, k = k1 + A*k2):
//   forward : thread n2 (<B) runs an A-point FFT over n1, multiplies by W_N^{n2 k1}, writes smem[k1][n2];
//             thread k1 (<A) runs a B-point FFT over n2 and ends up holding X[k1 + A*k2], k2 = 0..B-1.
//   inverse : the mirror image: thread k1 starts from X[k1 + A*k2] in registers, B-point inverse,
//             twiddle, smem[k1][n2]; thread n2 runs the A-point inverse and holds x[n1*B + n2].
// so a forward transform followed by an inverse one (the fused z pass) never needs a reordering.
//
// Shared-memory layout: element (k1, n2) of a line lives at base + (k1*BP + n2)*LS with BP = B|1 (odd
// pitch => both the row-wise writes and the column-wise reads are bank-conflict free for 8-byte
// accesses), LS = lane stride (T interleaved lines for the strided passes, 1 for the x passes).
//
// Every kernel is written as a sequence of `phase<k>` functions separated by block barriers.  The
// phases are __host__ __device__: tests/emu runs the very same code on the CPU, thread by thread.
#pragma once
#include "fft_defs.cuh"
#include "regfft_gen.cuh"              // RegFFT : scalar FADD / FFMA butterflies
#include "regfft_gen_packed.cuh"       // RegFFTP: the same dataflow on Blackwell's packed FP32x2 pipe (FADD2 / FFMA2)

#ifndef MVSIM_PACKED_FFT
#define MVSIM_PACKED_FFT 1
#endif

namespace mvsim {

// Measured on B200 (config 3): the packed butterflies cut the strided passes by 11-12 % (fewer issue slots, shorter
// dependent chains) but slow the x passes by 15 % (80-register cap at 3 CTAs/SM, aligned register pairs), so the x passes
// stay scalar.  MVSIM_PACKED_FFT=0 builds everything scalar (A/B measurements).
template <int N, int DIR, bool PK> struct RegSel { static MVSIM_HD void run(float2 (&x)[N]) { RegFFT<N, DIR>::run(x); } };
template <int N, int DIR> struct RegSel<N, DIR, true> { static MVSIM_HD void run(float2 (&x)[N]) { RegFFTP<N, DIR>::run(x); } };
constexpr bool kPackedStrided = MVSIM_PACKED_FFT != 0;
#ifndef MVSIM_PACKED_X
#define MVSIM_PACKED_X 0
#endif
// x passes on the packed pipe too (A/B builds, MVSIM_PACKED_X=1 python build.py -> libmvsim_px.so).  Measured again in round 2 with the
// 6-row forward CTAs (96 registers): forward 0.918 -> 0.967 ms (264 bytes of spills), inverse 0.252 -> 0.249 ms: stays off
constexpr bool kPackedX = MVSIM_PACKED_FFT != 0 && MVSIM_PACKED_X != 0;
// forward transform of x[0..K) with x[K..N) == 0 (not read), K = RegFFTPZ<N>::K: generated with the zero terms removed
template <int N, bool PK> struct RegSelZ {
    static constexpr int K = RegFFTPZ<N>::K;
    static MVSIM_HD void run(float2 (&x)[N])
    {
        MVSIM_UNROLL
        for (int i = K; i < N; ++i) x[i] = make_float2(0.f, 0.f);
        RegFFT<N, -1>::run(x);
    }
};
template <int N> struct RegSelZ<N, true> {
    static constexpr int K = RegFFTPZ<N>::K;
    static MVSIM_HD void run(float2 (&x)[N]) { RegFFTPZ<N>::run(x); }
};

template <int A_, int B_> struct LineShape {
    static constexpr int A = A_, B = B_;
    static constexpr bool IS_X = false;
    static constexpr int N = A * B;
    static constexpr int P = A > B ? A : B;   // threads per line
    static constexpr int BP = B | 1;          // odd pitch
    static constexpr int ELEMS = A * BP;      // float2 per line in shared memory
};

// ---- forward halves ---------------------------------------------------------------------------
// thread p < B.  x[n1] = line[p + n1*B] on entry.
template <int A, int B, bool PK = false> MVSIM_HD void fwd_first(int p, float2 (&x)[A], float2* sm, int base, int ls, const float2* __restrict__ tw)
{
    constexpr int BP = B | 1;
    RegSel<A, -1, PK>::run(x);
    MVSIM_UNROLL
    for (int k1 = 0; k1 < A; ++k1) {
        const float2 v = k1 == 0 ? x[0] : cmul(x[k1], tw[k1 * p]);
        sm[base + (k1 * BP + p) * ls] = v;
    }
}
// the same for a zero-extended line of which only x[0..K) carries data, K = RegSelZ<A, PK>::K
template <int A, int B, bool PK = false> MVSIM_HD void fwd_first_zext(int p, float2 (&x)[A], float2* sm, int base, int ls, const float2* __restrict__ tw)
{
    constexpr int BP = B | 1;
    RegSelZ<A, PK>::run(x);
    MVSIM_UNROLL
    for (int k1 = 0; k1 < A; ++k1) {
        const float2 v = k1 == 0 ? x[0] : cmul(x[k1], tw[k1 * p]);
        sm[base + (k1 * BP + p) * ls] = v;
    }
}
// thread p < A.  On exit y[k2] = X[p + A*k2].
template <int A, int B, bool PK = false> MVSIM_HD void fwd_second(int p, float2 (&y)[B], const float2* sm, int base, int ls)
{
    constexpr int BP = B | 1;
    MVSIM_UNROLL
    for (int n2 = 0; n2 < B; ++n2) y[n2] = sm[base + (p * BP + n2) * ls];
    RegSel<B, -1, PK>::run(y);
}

// ---- inverse halves (unscaled) ----------------------------------------------------------------
// thread p < A.  y[k2] = X[p + A*k2] on entry.
template <int A, int B, bool PK = false> MVSIM_HD void inv_first(int p, float2 (&y)[B], float2* sm, int base, int ls, const float2* __restrict__ tw)
{
    constexpr int BP = B | 1;
    RegSel<B, 1, PK>::run(y);
    MVSIM_UNROLL
    for (int n2 = 0; n2 < B; ++n2) {
        const float2 v = n2 == 0 ? y[0] : cmulc(y[n2], tw[n2 * p]);
        sm[base + (p * BP + n2) * ls] = v;
    }
}
// thread p < B.  On exit x[n1] = line[p + n1*B].
template <int A, int B, bool PK = false> MVSIM_HD void inv_second(int p, float2 (&x)[A], const float2* sm, int base, int ls)
{
    constexpr int BP = B | 1;
    MVSIM_UNROLL
    for (int k1 = 0; k1 < A; ++k1) x[k1] = sm[base + (k1 * BP + p) * ls];
    RegSel<A, 1, PK>::run(x);
}

// Element `idx` of a strided line through its 32-bit element stride e (host side guarantees n * e < 2^28, so the byte
// stride fits 32 bits too): ONE widening multiply-add (IMAD.WIDE.U32) instead of multiply + 64-bit add + scaled 64-bit add.
MVSIM_HD const float2* at32(const float2* base, unsigned idx, unsigned e)
{
    return reinterpret_cast<const float2*>(reinterpret_cast<const char*>(base) + (unsigned long long)idx * (unsigned long long)(e * 8u));
}
MVSIM_HD float2* at32(float2* base, unsigned idx, unsigned e)
{
    return reinterpret_cast<float2*>(reinterpret_cast<char*>(base) + (unsigned long long)idx * (unsigned long long)(e * 8u));
}

// Gathers x[n1] = ext(line)[p + n1*B - left], n1 < A, from a strided line.  All index arithmetic is done
// before the first load so the A loads of a thread are in flight together.
// e32 != 0: every element offset of the line (index * estride) fits 32 bits, so one 32-bit multiply + one widening
// address add per element replaces the 64-bit multiply (the integer pipe is ~40 % of these kernels' instructions).
template <int A, int B> MVSIM_HD void gather_line(float2 (&x)[A], const float2* __restrict__ src, long long estride, unsigned e32, int p, int left,
                                                 int n_src, int ext)
{
    // `left` already includes the block offset of overlap-save blocks (padded index q holds source q - left)
    if (ext == EXT_MIRROR1) {
        int idx[A];
        MVSIM_UNROLL
        for (int n1 = 0; n1 < A; ++n1) idx[n1] = mirror_once(p + n1 * B - left, n_src);
        if (e32) {
            MVSIM_UNROLL
            for (int n1 = 0; n1 < A; ++n1) x[n1] = *at32(src, (unsigned)idx[n1], e32);
        } else {
            MVSIM_UNROLL
            for (int n1 = 0; n1 < A; ++n1) x[n1] = src[idx[n1] * estride];
        }
    } else if (ext == EXT_ZERO) {
        if (e32) {
            MVSIM_UNROLL
            for (int n1 = 0; n1 < A; ++n1) {
                const int n = p + n1 * B - left;
                const bool ok = (unsigned)n < (unsigned)n_src;
                const float2 v = *at32(src, (unsigned)(ok ? n : 0), e32);      // clamped index + select keeps the loads batched
                x[n1] = ok ? v : make_float2(0.f, 0.f);
            }
        } else {
            MVSIM_UNROLL
            for (int n1 = 0; n1 < A; ++n1) {
                const int n = p + n1 * B - left;
                const bool ok = (unsigned)n < (unsigned)n_src;
                const float2 v = src[(ok ? n : 0) * estride];
                x[n1] = ok ? v : make_float2(0.f, 0.f);
            }
        }
    } else {
        MVSIM_UNROLL
        for (int n1 = 0; n1 < A; ++n1) x[n1] = src[mirror_single(p + n1 * B - left, n_src) * estride];
    }
}

// ==============================================================================================
// Strided forward pass (y pass, and the z pass of the PSF spectrum).  A CTA owns T neighbouring
// kx columns (contiguous in memory) of one `outer` slab and transforms them along the strided axis.
// The source line (length n_src) is extended on the fly: mirror-single (image; imglib2
// Views.extendMirrorSingle) or zero (kernel; Views.extendValue(kernel, 0)); `left` = kdim-1-kdim/2.
// ==============================================================================================
constexpr int kMaxRanks = 16;

struct StridedParams {
    const float2* in;
    float2* out;
    const float2* tw;       // exp(-2 pi i m / N), m < N
    int kx_count;           // complex columns per row
    int n_src;              // valid source samples along the line
    int left;               // padded index p holds source p - left
    int ext;                // EXT_MIRROR1 / EXT_ZERO / EXT_MIRROR_GENERAL (fft_defs.cuh)
    int crop0, n_out;       // inverse: store padded indices [crop0, crop0 + n_out)
    long long in_estride, in_ostride, out_estride, out_ostride;   // in float2 units
    unsigned in_e32, out_e32;   // != 0: the element stride again, when every offset inside a line fits 32 bits (set by strided_fill_e32)
    long long in_tstride, out_tstride;  // stride between kx tiles: T for row-major [..][KXc], Z*N*T for tile-major [KT][..][..][T]
    int swap_grid;          // 0: blockIdx.x = kx tile, .y = outer; 1: blockIdx.x = outer (neighbouring CTAs share DRAM pages)
    int tile0;              // global index of this launch's tile 0 (slab decomposition: a rank owns tiles [tile0, tile0 + n))
    int in_tile_global, out_tile_global;   // 1: that side is indexed by the global tile (row-major U1/P1), 0: by the local tile
    int out_offset;         // inverse: cropped sample o is stored at line index o + out_offset (overlap-save blocks along y)
    float scale;            // forward: multiplied into the output (folds 1/N and PSF scaling)
    // Slab decomposition with peer-to-peer stores (forward y pass only, n_peers > 1): the kx tile `t` of this rank's planes
    // goes straight into the z-pass buffer of its owner d = t / peer_tiles, segment my_rank:
    //   out_peers[d] + ((my_rank * peer_tiles + t - d * peer_tiles) * out_tstride) + outer * out_ostride + k * out_estride
    // (out_tstride = Zl*Ny*T).  The transfer over NVLink overlaps the transform tile by tile; no all-to-all pass exists.
    int n_peers, my_rank, peer_tiles;
    unsigned peer_tiles_magic;
    float2* out_peers[kMaxRanks];
    // L2 prefetch (forward image pass on row-major input, one GPU): thread 0 of CTA i asks the TMA unit for the source columns of
    // CTA i + prefetch_dist (launch order: kx tile fastest).  0 = off.
    int prefetch_dist, grid_x, grid_y;
    alignas(64) unsigned long long in_tmap[16];   // CUtensorMap over `in` as float32 [outer][n_src][2*kx_count], box [1][256][2T]
};
constexpr int kPrefetchBoxRows = 256;

struct NoState {};

// host side: enable the 32-bit element offsets of a strided pass over lines of padded length n when they cannot overflow
inline void strided_fill_e32(StridedParams& q, int n)
{
    long long top = n;
    if (q.n_src > top) top = q.n_src;
    if ((long long)q.n_out + q.out_offset > top) top = (long long)q.n_out + q.out_offset;
    q.in_e32 = (q.in_estride > 0 && top * q.in_estride < 0x0fffffffLL) ? (unsigned)q.in_estride : 0u;
    q.out_e32 = (q.out_estride > 0 && top * q.out_estride < 0x0fffffffLL) ? (unsigned)q.out_estride : 0u;
}

template <int A_, int B_, int T_> struct StridedFwd : LineShape<A_, B_> {
    using S = LineShape<A_, B_>;
    static constexpr int A = A_, B = B_, T = T_;
    static constexpr int THREADS = T * S::P;
    static constexpr int SMEM_BYTES = S::ELEMS * T * (int)sizeof(float2);
    static constexpr int NPH = 2;
    using Params = StridedParams;
    using State = NoState;

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int by, int tid, float2* sm, State&)
    {
        const int lane = tid % T, p = tid / T;
        const int tile = q.swap_grid ? by : bx, outer = q.swap_grid ? bx : by;
        const int tin = q.in_tile_global ? tile + q.tile0 : tile, tout = q.out_tile_global ? tile + q.tile0 : tile;
        const bool active = (tile + q.tile0) * T + lane < q.kx_count;
        if (PH == 0) {
#ifdef __CUDA_ARCH__
            if (tid == 0 && q.prefetch_dist > 0) {
                const int lin = by * q.grid_x + bx + q.prefetch_dist;
                const int o2 = lin / q.grid_x, t2 = lin - o2 * q.grid_x;
                if (o2 < q.grid_y)
                {
                    // source rows this launch touches: padded index i holds source i - left (overlap-save blocks use a window)
                    const int r0 = q.left < 0 ? -q.left : 0;
                    const int r1 = S::N - q.left < q.n_src ? S::N - q.left : q.n_src;
                    for (int r = r0; r < r1; r += kPrefetchBoxRows) tma_prefetch_3d(q.in_tmap, 2 * T * t2, r, o2);
                }
            }
#endif
            if (p < B && active) {
                float2 x[A];
                const float2* src = q.in + tin * q.in_tstride + outer * q.in_ostride + lane;
                constexpr int K = RegSelZ<A, kPackedStrided>::K;
                if (q.ext == EXT_ZERO && q.left == 0 && q.n_src <= K * B) {
                    // zero-extended line no longer than a fifth of the padded length (the PSF's y pass: 128 of 1152 samples):
                    // only x[0..K) carry data, first half with the zero terms removed at generation time
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < K; ++n1) {
                        const int n = p + n1 * B;
                        const bool ok = n < q.n_src;
                        const float2 v = src[(long long)(ok ? n : 0) * q.in_estride];
                        x[n1] = ok ? v : make_float2(0.f, 0.f);
                    }
                    fwd_first_zext<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                } else {
                    gather_line<A, B>(x, src, q.in_estride, q.in_e32, p, q.left, q.n_src, q.ext);
                    fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                }
            }
        } else {
            if (p < A && active) {
                float2 y[B];
                fwd_second<A, B, kPackedStrided>(p, y, sm, lane, T);
                float2* dst;
                if (q.n_peers > 1) {
                    const int d = q.peer_tiles == 1 ? tout : (int)umulhi32((uint32_t)tout, q.peer_tiles_magic);
                    dst = q.out_peers[d] + (long long)(q.my_rank * q.peer_tiles + tout - d * q.peer_tiles) * q.out_tstride + outer * q.out_ostride + lane;
                } else {
                    dst = q.out + tout * q.out_tstride + outer * q.out_ostride + lane;
                }
                if (q.out_e32) {
                    MVSIM_UNROLL
                    for (int k2 = 0; k2 < B; ++k2)
                        *at32(dst, (unsigned)(p + A * k2), q.out_e32) = make_float2(y[k2].x * q.scale, y[k2].y * q.scale);
                } else {
                    MVSIM_UNROLL
                    for (int k2 = 0; k2 < B; ++k2)
                        dst[(p + A * k2) * q.out_estride] = make_float2(y[k2].x * q.scale, y[k2].y * q.scale);
                }
            }
        }
    }
};

// Strided inverse pass (y inverse): natural-order spectrum in, padded indices [crop0, crop0+n_out) out.
template <int A_, int B_, int T_> struct StridedInv : LineShape<A_, B_> {
    using S = LineShape<A_, B_>;
    static constexpr int A = A_, B = B_, T = T_;
    static constexpr int THREADS = T * S::P;
    static constexpr int SMEM_BYTES = S::ELEMS * T * (int)sizeof(float2);
    static constexpr int NPH = 2;
    using Params = StridedParams;
    using State = NoState;

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int by, int tid, float2* sm, State&)
    {
        const int lane = tid % T, p = tid / T;
        const int tile = q.swap_grid ? by : bx, outer = q.swap_grid ? bx : by;
        const int tin = q.in_tile_global ? tile + q.tile0 : tile, tout = q.out_tile_global ? tile + q.tile0 : tile;
        const bool active = (tile + q.tile0) * T + lane < q.kx_count;
        if (PH == 0) {
#ifdef __CUDA_ARCH__
            if (tid == 0 && q.prefetch_dist > 0) {
                // tile-major input: the T columns x N rows of a CTA are one contiguous range
                const int lin = by * q.grid_x + bx + q.prefetch_dist;
                const int o2 = lin / q.grid_x, t2 = lin - o2 * q.grid_x;
                if (o2 < q.grid_y) bulk_prefetch_l2(q.in + t2 * q.in_tstride + o2 * q.in_ostride, (unsigned)(S::N * T * sizeof(float2)));
            }
#endif
            if (p < A && active) {
                float2 y[B];
                const float2* src = q.in + tin * q.in_tstride + outer * q.in_ostride + lane;
                if (q.in_e32) {
                    MVSIM_UNROLL
                    for (int k2 = 0; k2 < B; ++k2) y[k2] = *at32(src, (unsigned)(p + A * k2), q.in_e32);
                } else {
                    MVSIM_UNROLL
                    for (int k2 = 0; k2 < B; ++k2) y[k2] = src[(p + A * k2) * q.in_estride];
                }
                inv_first<A, B, kPackedStrided>(p, y, sm, lane, T, q.tw);
            }
        } else {
            if (p < B && active) {
                float2 x[A];
                inv_second<A, B, kPackedStrided>(p, x, sm, lane, T);
                float2* dst = q.out + tout * q.out_tstride + outer * q.out_ostride + lane;
                if (q.out_e32) {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        const int o = p + n1 * B - q.crop0;
                        if ((unsigned)o < (unsigned)q.n_out) *at32(dst, (unsigned)(o + q.out_offset), q.out_e32) = x[n1];
                    }
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        const int o = p + n1 * B - q.crop0;
                        if ((unsigned)o < (unsigned)q.n_out) dst[(o + q.out_offset) * q.out_estride] = x[n1];
                    }
                }
            }
        }
    }
};

// ==============================================================================================
// Fused z pass: mirror-extended forward FFT along z, multiply by the PSF spectrum H (the complex
// multiply of FFTConvolution.multiplyComplex), inverse FFT along z, store only the cropped range.
// In place: a CTA reads its T lines completely before it writes them.
// ==============================================================================================
struct ZFusedParams {
    float2* u;              // [S][tiles][zg][Ny][T] in place: S segments of zg planes (S = ranks of a slab-decomposed run: the
                            // all-to-all receive layout; S = 1, zg = Z on one GPU, i.e. plain tile-major [KT][Z][Ny][T])
    const float2* h;        // tile-major [tiles][Nz][Ny][T], already scaled by 1/(N*Ny*Nz)
    const float2* tw;
    int kx_count, n_src, left, crop0;
    int ext;                // EXT_MIRROR1 or EXT_MIRROR_GENERAL
    int tile0;              // global index of local tile 0
    int zg;                 // planes per segment
    unsigned zg_magic;      // div_magic(zg)
    unsigned keep_magic;    // div_magic(keep_inc)
    int keep_inc, n_keep;   // keep_inc > 1 (single segment only): store only z = 0, inc, 2 inc, ... compacted to planes
                            // 0..n_keep-1 and the SUM of all other cropped z in plane n_keep (enough for extractSlices + the
                            // mean of adjustImage).  ZFusedPoly stores the sum of ALL cropped z there instead (conv_middle_z
                            // reports which; the inverse x pass then counts that plane alone, XParams::sum_row0)
    long long estride;      // z stride inside a segment (= Ny*T), also the kz stride of h
    int estride32;          // != 0: one segment (single GPU) and every in-line offset z*estride fits 32 bits: the loaders and storers
                            // then spend one 32-bit multiply per element instead of the segmented 64-bit address arithmetic
    long long ostride;      // ky stride (= T), same for u and h
    long long u_tstride;    // kx-tile stride of u inside a segment (= zg*Ny*T)
    long long seg_stride;   // segment stride of u (= tiles*zg*Ny*T)
    long long h_tstride;    // kx-tile stride of h (= Nz*Ny*T)
    // Slab decomposition with peer-to-peer stores (n_peers > 1): the planes [seg zg, (seg + 1) zg) of a line go to their owner seg as
    // ONE bulk copy shared -> peer global: out_peers[seg] + (((tile0 + tile) * Ny + ky) * zg) * T, i.e. the inverse-side buffers are
    // laid out [KT][Ny][zg][T] in this mode (Ny = estride / T); the all-to-all back is fused into the store.
    int n_peers;
    float2* out_peers[kMaxRanks];
    // h_mode 1 (ZFusedOTF): the PSF spectrum is never materialised -- the kernel transforms the PSF's partial spectrum
    // p2 = [tiles][KZ][Ny][T] (x and y transformed, z still spatial, zero extended, pre-scaled) along z itself and keeps the
    // line's H values in shared memory.  Saves writing and re-reading Nz/KZ times more data per convolution.
    int h_mode;
    const float2* p2;
    int k_src;              // KZ: valid z samples of the PSF
    long long p2_tstride;   // kx-tile stride of p2 (= KZ*Ny*T)
    int use_tma;            // 1: the H tile is fetched by the TMA unit through h_tmap (device only), 0: cp.async per thread
    // L2 prefetch (ZFusedOTF, one GPU): thread 0 of CTA i asks the TMA unit to bring the image tile of CTA i + prefetch_dist
    // (launch order: ky fastest) into the L2, so that CTA's gather finds its lines there instead of in DRAM.  0 = off.
    int prefetch_dist, grid_x, grid_y;
    // ZFusedDec: D[k] = sum over the cropped range of exp(+2 pi i n k / N): sum of the cropped outputs = sum_k Yhat[k] D[k]
    const float2* dtab;
    alignas(64) unsigned long long u_tmap[16];   // CUtensorMap over u as float32 [tiles][Z][Ny][2T], box [1][128][1][2T]
    alignas(64) unsigned long long h_tmap[16];   // CUtensorMap over h (or, h_mode 1, over p2) as float32 [tiles][rows][Ny][2T], box [1][128][1][2T]
};

constexpr int kTmaBoxRows = 128;    // kz rows per TMA box

// offset of global plane z in the segmented layout
MVSIM_HD long long zfused_plane_offset(const ZFusedParams& q, int z)
{
    const int seg = q.zg == 1 ? z : (int)umulhi32((uint32_t)z, q.zg_magic);   // div_magic(1) does not fit 32 bits
    return seg * q.seg_stride + (z - seg * q.zg) * q.estride;
}

template <int B> struct RegState { float2 y[B]; };

template <int A_, int B_, int T_> struct ZFused : LineShape<A_, B_> {
    using S = LineShape<A_, B_>;
    static constexpr int A = A_, B = B_, T = T_;
    static constexpr int THREADS = T * S::P;
    // exchange area (rounded to 128 B) + the H tile [N rounded up to whole TMA boxes][T] (fetched asynchronously -- TMA or
    // cp.async -- while the forward transform runs) + one mbarrier
    static constexpr int EXCH_ELEMS = (S::ELEMS * T + 15) / 16 * 16;
    static constexpr int H_ROWS = (S::N + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows;
    static constexpr int SMEM_BYTES = (EXCH_ELEMS + H_ROWS * T) * (int)sizeof(float2) + 16;
    static constexpr int NPH = 6;
    using Params = ZFusedParams;
    using State = RegState<B>;

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int by, int tid, float2* sm, State& st)
    {
        // blockIdx.x = ky (CTAs that run together touch neighbouring 64-byte chunks of every z plane),
        // blockIdx.y = kx tile
        const int lane = tid % T, p = tid / T;
        const int tile = by, outer = bx;
        const bool active = (tile + q.tile0) * T + lane < q.kx_count;
        float2* smh = sm + EXCH_ELEMS;
#ifdef __CUDA_ARCH__
        uint64_t* bar = reinterpret_cast<uint64_t*>(smh + H_ROWS * T);
#endif
        if (PH == 0) {
            if (q.h_mode != 0) {
                // H values were computed into smh by ZFusedOTF's PSF phases
            } else
#ifdef __CUDA_ARCH__
            if (q.use_tma) {
                // one thread programs the TMA unit: ceil(N/128) boxes of 128 kz rows x 64 bytes, completion on the mbarrier
                if (tid == 0) {
                    constexpr int NBOX = H_ROWS / kTmaBoxRows;
                    mbar_init(bar, 1);
                    mbar_expect_tx(bar, (unsigned)(NBOX * kTmaBoxRows * T * sizeof(float2)));
                    MVSIM_UNROLL
                    for (int b = 0; b < NBOX; ++b)
                        tma_load_4d(smh + b * kTmaBoxRows * T, q.h_tmap, 0, outer, b * kTmaBoxRows, tile, bar);
                }
            } else
#endif
            {   // H tile -> shared memory, 16 bytes per copy, rows of T float2
                constexpr int CH = T / 2;       // 16-byte chunks per row
                const float2* hs = q.h + tile * q.h_tstride + outer * q.ostride;
                for (int c = tid; c < S::N * CH; c += THREADS) {
                    const int row = c / CH, part = c % CH;
                    cp_async16(smh + row * T + part * 2, hs + row * q.estride + part * 2);
                }
            }
            if (p < B && active) {
                float2 x[A];
                const float2* __restrict__ src = q.u + tile * q.u_tstride + outer * q.ostride + lane;
                int idx[A];
                if (q.ext == EXT_MIRROR1) {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) idx[n1] = mirror_once(p + n1 * B - q.left, q.n_src);
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) idx[n1] = mirror_single(p + n1 * B - q.left, q.n_src);
                }
                if (q.estride32) {
                    const unsigned e = (unsigned)q.estride32;
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) x[n1] = *at32(src, (unsigned)idx[n1], e);
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) x[n1] = src[zfused_plane_offset(q, idx[n1])];
                }
                fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
            }
            cp_async_wait_all();
        } else if (PH == 1) {
            if (p < A && active) {
                fwd_second<A, B, kPackedStrided>(p, st.y, sm, lane, T);
#ifdef __CUDA_ARCH__
                if (q.use_tma) mbar_wait(bar, 0);       // the barrier before this phase made the init visible
#endif
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) st.y[k2] = cmul(st.y[k2], smh[(p + A * k2) * T + lane]);
            }
        } else if (PH == 2) {
            if (p < A && active) inv_first<A, B, kPackedStrided>(p, st.y, sm, lane, T, q.tw);
        } else if (PH == 3) {
            if (p < B && active) {
                float2 x[A];
                inv_second<A, B, kPackedStrided>(p, x, sm, lane, T);
                float2* dst = q.u + tile * q.u_tstride + outer * q.ostride + lane;
                if (q.keep_inc > 1) {
                    // whole-view call: the line goes to the (now free) H area in natural order; the next phase picks the kept
                    // planes and sums the rest with a fixed trip count per thread (the per-element keep test in registers cost
                    // 20 instructions per output: a sixth of the kernel)
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) smh[(p + n1 * B) * T + lane] = x[n1];
                } else if (q.n_peers > 1) {
                    // slab decomposition, exchange fused into the store: the line goes to the H area in natural order and the
                    // next phase ships each owner's z range with ONE bulk copy (8 KB at config 5 on 8 GPUs) instead of 64-byte
                    // peer stores (measured in round 1: 6.2 ms against 3.9 ms with local stores)
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) smh[(p + n1 * B) * T + lane] = x[n1];
#ifdef __CUDA_ARCH__
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> visible to the bulk copy
#endif
                } else {
                    if (q.estride32) {
                        const unsigned e = (unsigned)q.estride32;
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < A; ++n1) {
                            const int o = p + n1 * B - q.crop0;
                            if ((unsigned)o < (unsigned)q.n_src) *at32(dst, (unsigned)o, e) = x[n1];
                        }
                    } else {
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < A; ++n1) {
                            const int o = p + n1 * B - q.crop0;
                            if ((unsigned)o < (unsigned)q.n_src) dst[zfused_plane_offset(q, o)] = x[n1];
                        }
                    }
                }
            }
        } else if (PH == 4) {
            // (barrier before: the line is complete in the H area and nobody reads the exchange area any more)
            if (q.n_peers > 1) {
                // owner `seg` of the planes [seg zg, (seg + 1) zg) receives them as one contiguous run of its inverse-side buffer,
                // laid out [KT][Ny][zg][T] in this mode (z fastest after the lanes): dst = peer + ((tile, ky) zg) T
                if (tid < q.n_peers) {
                    const int seg = tid;
                    const float2* srcp = smh + (long long)(q.crop0 + seg * q.zg) * T;
                    float2* dstp = q.out_peers[seg] + ((long long)(q.tile0 + tile) * (q.estride / T) + outer) * q.zg * T;      // estride / T = Ny
                    const unsigned bytes = (unsigned)(q.zg * T * sizeof(float2));
#ifdef __CUDA_ARCH__
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                                 ::"l"(dstp), "r"((unsigned)__cvta_generic_to_shared(srcp)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the CTA's shared memory outlives the read
#else
                    for (unsigned i = 0; i < bytes / sizeof(float2); ++i) dstp[i] = srcp[i];
#endif
                }
            }
            // thread (p, lane): sum of the cropped rows p, p+P, ... minus the kept ones among kz = p, p+P, ..., which it stores
            if (q.keep_inc > 1) {
                float2 acc = make_float2(0.f, 0.f);
                if (active) {
                    const float2* row = smh + q.crop0 * T + lane;
                    MVSIM_UNROLL4
                    for (int j = p; j < q.n_src; j += S::P) { const float2 v = row[j * T]; acc.x += v.x; acc.y += v.y; }
                    float2* dst = q.u + tile * q.u_tstride + outer * q.ostride + lane;
                    const int kstep = q.keep_inc * T;
                    if (q.estride32) {
                        const unsigned e = (unsigned)q.estride32;
                        for (int kz = p; kz < q.n_keep; kz += S::P) {
                            const float2 v = row[kz * kstep];
                            *at32(dst, (unsigned)kz, e) = v;
                            acc.x -= v.x; acc.y -= v.y;
                        }
                    } else {
                        for (int kz = p; kz < q.n_keep; kz += S::P) {
                            const float2 v = row[kz * kstep];
                            dst[kz * q.estride] = v;
                            acc.x -= v.x; acc.y -= v.y;
                        }
                    }
                }
                sm[p * T + lane] = acc;
            }
        } else {
            if (q.keep_inc > 1 && p == 0 && active) {
                float2 s = make_float2(0.f, 0.f);
                for (int j = 0; j < S::P; ++j) { s.x += sm[j * T + lane].x; s.y += sm[j * T + lane].y; }
                q.u[tile * q.u_tstride + outer * q.ostride + lane + q.n_keep * q.estride] = s;
            }
        }
    }
};

// Shared memory of ZFusedOTF (host side needs it without the template): exchange area + H area + 128 B for the mbarrier,
// plus, on the TMA path, the PSF tile.
constexpr int kSmemLimit = 227 * 1024;      // per CTA, sm_100
constexpr int zfused_otf_smem_base(int a, int b, int t)
{
    return ((a * (b | 1) * t + 15) / 16 * 16 + (a * b + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows * t + 16) * (int)sizeof(float2);
}
constexpr int zfused_otf_psf_tile_bytes(int k_src, int t) { return (k_src + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows * t * (int)sizeof(float2); }
// the TMA-fed PSF tile is used when it fits and does not cost a resident CTA (228 KB per SM, 1 KB reserved per CTA)
inline bool zfused_otf_tma_fits(int a, int b, int t, int k_src)
{
    const int base = zfused_otf_smem_base(a, b, t), with = base + zfused_otf_psf_tile_bytes(k_src, t);
    if (with > kSmemLimit) return false;
    const int c0 = 233472 / (base + 1024), c1 = 233472 / (with + 1024);
    return c1 >= (c0 < 4 ? c0 : 4);
}

// Fused z pass with the PSF spectrum computed on the fly (h_mode 1).  Phase order: the IMAGE line is transformed first and its
// spectrum parked in the H area, then the PSF line (P2, zero extended) is transformed and multiplied in, then ZFused's inverse
// half.  The PSF tile is requested from the TMA unit by one thread at CTA start and lands in its own small area while the
// image phases run, so only one DRAM round trip (the image gather) is exposed per CTA instead of two.
template <int A_, int B_, int T_> struct ZFusedOTF : ZFused<A_, B_, T_> {
    using Z = ZFused<A_, B_, T_>;
    using S = LineShape<A_, B_>;
    static constexpr int A = A_, B = B_, T = T_;
    static constexpr int NPH = 8;
    using Params = ZFusedParams;
    using State = typename Z::State;
    // [exchange][H area: parked image spectrum][mbarrier, padded to 128 B][PSF tile: ceil(KZ/128) TMA boxes of 128 rows x T]
    static constexpr int BAR_ELEMS = Z::EXCH_ELEMS + Z::H_ROWS * T;
    static constexpr int PSF_ELEMS0 = BAR_ELEMS + 16;
    static constexpr int SMEM_BYTES = PSF_ELEMS0 * (int)sizeof(float2);         // without the PSF tile (per-thread loads)
    static int smem_bytes(const Params& q) { return SMEM_BYTES + (q.use_tma ? zfused_otf_psf_tile_bytes(q.k_src, T) : 0); }
    static int smem_bytes_max() { const int m = SMEM_BYTES + Z::H_ROWS * T * (int)sizeof(float2); return m < kSmemLimit ? m : kSmemLimit; }
    static_assert(SMEM_BYTES == zfused_otf_smem_base(A_, B_, T_), "host-side shared memory formula out of sync");

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int by, int tid, float2* sm, State& st)
    {
        if (PH >= 3) {
            // multiply (the registers hold the PSF line's spectrum, the H area the image line's), inverse half, stores
            Z::template phase<(PH >= 3 ? PH - 2 : 1)>(q, bx, by, tid, sm, st);
            return;
        }
        const int lane = tid % T, p = tid / T;
        const int tile = by, outer = bx;
        const bool active = (tile + q.tile0) * T + lane < q.kx_count;
        float2* smh = sm + Z::EXCH_ELEMS;
        if (PH == 0) {
#ifdef __CUDA_ARCH__
            if (q.use_tma && tid == 0) {
                // one thread programs ceil(KZ/128) box loads (128 kz rows x 64 bytes each; rows beyond KZ are zero filled by the
                // unit); the tile is first read two barriers later
                uint64_t* bar = reinterpret_cast<uint64_t*>(sm + BAR_ELEMS);
                const int nbox = (q.k_src + kTmaBoxRows - 1) / kTmaBoxRows;
                mbar_init(bar, 1);
                mbar_expect_tx(bar, (unsigned)(nbox * kTmaBoxRows * T * sizeof(float2)));
                for (int b = 0; b < nbox; ++b) tma_load_4d(sm + PSF_ELEMS0 + b * kTmaBoxRows * T, q.h_tmap, 0, outer, b * kTmaBoxRows, tile, bar);
                if (q.prefetch_dist > 0) {
                    const int lin = by * q.grid_x + bx + q.prefetch_dist;
                    const int t2 = lin / q.grid_x, o2 = lin - t2 * q.grid_x;
                    if (t2 < q.grid_y)
                        for (int z = 0; z < q.n_src; z += kTmaBoxRows) tma_prefetch_4d(q.u_tmap, 0, o2, z, t2);
                }
            }
#endif
            Z::template phase<0>(q, bx, by, tid, sm, st);       // image line: gather (mirror extension) + first half
        } else if (PH == 1) {
            if (p < A && active) {
                float2 y[B];
                fwd_second<A, B, kPackedStrided>(p, y, sm, lane, T);
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) smh[(p + A * k2) * T + lane] = y[k2];     // read back by the same thread only
            }
        } else {
#ifdef __CUDA_ARCH__
            if (q.use_tma) {
                mbar_wait(reinterpret_cast<uint64_t*>(sm + BAR_ELEMS), 0);      // (two barriers since the init)
                if (p < B && active) {
                    const float2* tilep = sm + PSF_ELEMS0;
                    const int rows = (q.k_src + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows;
                    float2 x[A];
                    constexpr int K = RegSelZ<A, kPackedStrided>::K;
                    if (q.k_src <= K * B) {
                        // the usual case (PSF no longer than a fifth of the padded line): only x[0..K) carry data
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < K; ++n1) {
                            const int n = p + n1 * B;
                            x[n1] = n < rows ? tilep[n * T + lane] : make_float2(0.f, 0.f);
                        }
                        fwd_first_zext<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                    } else {
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < A; ++n1) {
                            const int n = p + n1 * B;
                            x[n1] = n < rows ? tilep[n * T + lane] : make_float2(0.f, 0.f);
                        }
                        fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                    }
                }
            } else
#endif
            if (p < B && active) {
                float2 x[A];
                const float2* __restrict__ src = q.p2 + tile * q.p2_tstride + outer * q.ostride + lane;
                constexpr int K = RegSelZ<A, kPackedStrided>::K;
                if (q.k_src <= K * B) {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < K; ++n1) {
                        const int n = p + n1 * B;
                        const bool ok = n < q.k_src;
                        const float2 v = src[(ok ? n : 0) * q.estride];
                        x[n1] = ok ? v : make_float2(0.f, 0.f);
                    }
                    fwd_first_zext<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        const int n = p + n1 * B;
                        const bool ok = n < q.k_src;
                        const float2 v = src[(ok ? n : 0) * q.estride];      // clamped index + select keeps the loads batched
                        x[n1] = ok ? v : make_float2(0.f, 0.f);              // (predicated loads measured slower: 3.39 vs 3.21 ms)
                    }
                    fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                }
            }
        }
    }
};

// ==============================================================================================
// Fused z pass of the whole-view call with a DECIMATED inverse (default wherever zfused_dec_ok() holds; measured on B200,
// config 3: 2.49 -> 2.13 ms against ZFusedOTF, profiles/r02_experiments.txt; MVSIM_Z_KERNEL=2 switches back for A/B runs).
// extractSlices keeps z = 0, INC, 2 INC, ... (:206), i.e. the
// padded outputs n = crop0 + INC kz.  With the split n = n1 B + n2 and INC | B these are exactly the columns
// n2 = r (mod INC), r = crop0 mod INC, of the exchange, so
//   - the first inverse half needs B/INC of its B outputs per thread (RegFFTPD: dead code removed by the generator),
//   - the second inverse half runs for B/INC of the B threads of a line,
//   - the plane that carries the sum of the dropped slices (adjustImage's mean) comes from ONE dot product: the sum of all
//     cropped outputs is sum_k Yhat[k] D[k] with D[k] = sum_{n in crop} exp(+2 pi i n k / N), minus the kept outputs.
// A_, B_ are the planner's (b, a): the first level has the larger transform here (config 3: A = 32, B = 20, INC = 5).
// Phases: image first half | image second half, park | PSF first half | PSF second half, multiply, dot with D' |
//         pruned inverse first half | second inverse half of the kept columns, SPLIT sub-transforms per column |
//         combine + stores | sum plane.
// The second inverse half used to run as ONE A-point transform per kept column on KEEP of the P threads of a line -- a single
// warp per CTA for ~600 dependent instructions while seven warps waited (ncu source view r02b: 15 % of all warp samples sat at
// that barrier, the phase was ~20 % of a CTA's lifetime).  It is now a radix-SPLIT step spread over SPLIT x as many threads:
// thread (column, q) transforms the inputs k1 = SPLIT k' + q (an A/SPLIT-point inverse), twiddles, parks G_q in the free H area;
// after a barrier thread (column, m[, half]) combines x[n' + (A/SPLIT) m] = sum_q G_q[n'] e^{+2 pi i m q / SPLIT} and stores.
// The plane that carries the sum of the DROPPED slices is one dot product with D'[k] = sum over the cropped, non-kept n of
// e^{+2 pi i n k / N} (host table): no bookkeeping of the kept outputs is needed.
constexpr bool dec_regfft_size(int m)
{
    return m == 2 || m == 3 || m == 4 || m == 5 || m == 6 || m == 8 || m == 9 || m == 10 || m == 12;
}
// radix of the split of the second inverse half: the largest of 4, 3, 5, 2 that divides a, leaves a generated sub-transform and
// fits the threads of a line (keep columns x split <= p); 1 = no split
constexpr int dec_split(int a, int keep, int p)
{
    const int cand[4] = { 4, 3, 5, 2 };
    for (int i = 0; i < 4; ++i)
        if (a % cand[i] == 0 && dec_regfft_size(a / cand[i]) && keep * cand[i] <= p) return cand[i];
    return 1;
}
// may the planner's split (a, b) of a z line run the decimated kernel ZFusedDec<b, a, T, inc>?
constexpr bool zfused_dec_ok(int a, int b, int inc) { return (inc == 3 || inc == 5) && a % inc == 0 && a / inc >= 2 && a + b <= a * b; }

template <int A_, int B_, int T_, int INC_> struct ZFusedDec : LineShape<A_, B_> {
    using S = LineShape<A_, B_>;
    static constexpr int A = A_, B = B_, T = T_, INC = INC_;
    static_assert(B_ % INC_ == 0 && B_ / INC_ >= 2, "the kept outputs must be whole columns of the exchange");
    static constexpr int KEEP = B / INC;          // kept columns per line
    static constexpr int THREADS = T * S::P;
    static constexpr int NPH = 8;
    static constexpr int SPLIT = dec_split(A_, B_ / INC_, S::P);    // radix of the split second inverse half
    static constexpr int M = A / SPLIT;                             // sub-transform length
    static constexpr int RS = (S::P / (KEEP * SPLIT) >= 2 && M % 2 == 0) ? 2 : 1;     // threads per (column, m) in the combine phase
    // H area after the multiply (float2 elements): [0, A T) dot partials | [G0, G0 + KEEP A T) the twiddled sub-transform outputs G_q
    static constexpr int G0 = A * T;
    static_assert(G0 + KEEP * A * T <= S::N * T, "G area must fit the parked-spectrum area");
    static constexpr int EXCH_ELEMS = (S::ELEMS * T + 15) / 16 * 16;
    static constexpr int H_ROWS = (S::N + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows;
    static constexpr int BAR_ELEMS = EXCH_ELEMS + H_ROWS * T;
    static constexpr int PSF_ELEMS0 = BAR_ELEMS + 16;
    static constexpr int SMEM_BYTES = PSF_ELEMS0 * (int)sizeof(float2);
    static_assert(SMEM_BYTES == zfused_otf_smem_base(A_, B_, T_), "host-side shared memory formula out of sync");
    using Params = ZFusedParams;
    using State = RegState<B>;
    static int smem_bytes(const Params& q) { return SMEM_BYTES + (q.use_tma ? zfused_otf_psf_tile_bytes(q.k_src, T) : 0); }
    static int smem_bytes_max() { const int m = SMEM_BYTES + H_ROWS * T * (int)sizeof(float2); return m < kSmemLimit ? m : kSmemLimit; }

    // first inverse half restricted to the columns n2 = R + INC j: twiddle and store
    template <int R> static MVSIM_HD void inv_first_kept(int p, const float2 (&y)[B], float2* sm, int lane, const float2* __restrict__ tw)
    {
        constexpr int BP = B | 1;
        float2 o[KEEP];
        RegFFTPD<B, INC, R>::run(y, o);
        MVSIM_UNROLL
        for (int j = 0; j < KEEP; ++j) {
            const int n2 = R + INC * j;
            const float2 v = n2 == 0 ? o[j] : cmulc(o[j], tw[n2 * p]);
            sm[lane + (p * BP + n2) * T] = v;
        }
    }

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int by, int tid, float2* sm, State& st)
    {
        const int lane = tid % T, p = tid / T;
        const int tile = by, outer = bx;
        const bool active = (tile + q.tile0) * T + lane < q.kx_count;
        float2* smh = sm + EXCH_ELEMS;
        const int r = q.crop0 % INC;
        if (PH == 0) {
#ifdef __CUDA_ARCH__
            if (q.use_tma && tid == 0) {
                uint64_t* bar = reinterpret_cast<uint64_t*>(sm + BAR_ELEMS);
                const int nbox = (q.k_src + kTmaBoxRows - 1) / kTmaBoxRows;
                mbar_init(bar, 1);
                mbar_expect_tx(bar, (unsigned)(nbox * kTmaBoxRows * T * sizeof(float2)));
                for (int b = 0; b < nbox; ++b) tma_load_4d(sm + PSF_ELEMS0 + b * kTmaBoxRows * T, q.h_tmap, 0, outer, b * kTmaBoxRows, tile, bar);
                if (q.prefetch_dist > 0) {
                    const int lin = by * q.grid_x + bx + q.prefetch_dist;
                    const int t2 = lin / q.grid_x, o2 = lin - t2 * q.grid_x;
                    if (t2 < q.grid_y)
                        for (int z = 0; z < q.n_src; z += kTmaBoxRows) tma_prefetch_4d(q.u_tmap, 0, o2, z, t2);
                }
            }
#endif
            if (p < B && active) {
                float2 x[A];
                const float2* __restrict__ src = q.u + tile * q.u_tstride + outer * q.ostride + lane;
                int idx[A];
                if (q.ext == EXT_MIRROR1) {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) idx[n1] = mirror_once(p + n1 * B - q.left, q.n_src);
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) idx[n1] = mirror_single(p + n1 * B - q.left, q.n_src);
                }
                const unsigned e = (unsigned)q.estride32;
                MVSIM_UNROLL
                for (int n1 = 0; n1 < A; ++n1) x[n1] = *at32(src, (unsigned)idx[n1], e);
                fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
            }
        } else if (PH == 1) {
            if (p < A && active) {
                float2 y[B];
                fwd_second<A, B, kPackedStrided>(p, y, sm, lane, T);
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) smh[(p + A * k2) * T + lane] = y[k2];     // read back by the same thread only
            }
        } else if (PH == 2) {
            constexpr int K = RegSelZ<A, kPackedStrided>::K;
            const bool pruned = q.k_src <= K * B;
#ifdef __CUDA_ARCH__
            if (q.use_tma) {
                mbar_wait(reinterpret_cast<uint64_t*>(sm + BAR_ELEMS), 0);
                if (p < B && active) {
                    const float2* tilep = sm + PSF_ELEMS0;
                    const int rows = (q.k_src + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows;
                    float2 x[A];
                    if (pruned) {
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < K; ++n1) {
                            const int n = p + n1 * B;
                            x[n1] = n < rows ? tilep[n * T + lane] : make_float2(0.f, 0.f);
                        }
                        fwd_first_zext<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                    } else {
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < A; ++n1) {
                            const int n = p + n1 * B;
                            x[n1] = n < rows ? tilep[n * T + lane] : make_float2(0.f, 0.f);
                        }
                        fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                    }
                }
            } else
#endif
            if (p < B && active) {
                float2 x[A];
                const float2* __restrict__ src = q.p2 + tile * q.p2_tstride + outer * q.ostride + lane;
                if (pruned) {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < K; ++n1) {
                        const int n = p + n1 * B;
                        const bool ok = n < q.k_src;
                        const float2 v = src[(ok ? n : 0) * q.estride];
                        x[n1] = ok ? v : make_float2(0.f, 0.f);
                    }
                    fwd_first_zext<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        const int n = p + n1 * B;
                        const bool ok = n < q.k_src;
                        const float2 v = src[(ok ? n : 0) * q.estride];
                        x[n1] = ok ? v : make_float2(0.f, 0.f);
                    }
                    fwd_first<A, B, kPackedStrided>(p, x, sm, lane, T, q.tw);
                }
            }
        } else if (PH == 3) {
            if (p < A && active) {
                fwd_second<A, B, kPackedStrided>(p, st.y, sm, lane, T);
                float2 acc = make_float2(0.f, 0.f);
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) {
                    st.y[k2] = cmul(st.y[k2], smh[(p + A * k2) * T + lane]);
                    const float2 d = q.dtab[p + A * k2];
                    acc.x += st.y[k2].x * d.x - st.y[k2].y * d.y;
                    acc.y += st.y[k2].x * d.y + st.y[k2].y * d.x;
                }
                smh[p * T + lane] = acc;        // own slot (k2 = 0), read in the last phase
            }
        } else if (PH == 4) {
            if (p < A && active) {
                // uniform switch on the residue class of the kept outputs
                if (INC == 3) {
                    if (r == 0) inv_first_kept<0>(p, st.y, sm, lane, q.tw);
                    else if (r == 1) inv_first_kept<1>(p, st.y, sm, lane, q.tw);
                    else inv_first_kept<2>(p, st.y, sm, lane, q.tw);
                } else {
                    if (r == 0) inv_first_kept<0>(p, st.y, sm, lane, q.tw);
                    else if (r == 1) inv_first_kept<1>(p, st.y, sm, lane, q.tw);
                    else if (r == 2) inv_first_kept<2>(p, st.y, sm, lane, q.tw);
                    else if (r == 3) inv_first_kept<(INC > 3 ? 3 : 0)>(p, st.y, sm, lane, q.tw);
                    else inv_first_kept<(INC > 4 ? 4 : 0)>(p, st.y, sm, lane, q.tw);
                }
            }
        } else if (PH == 5) {
            // second inverse half, part 1: thread (column j, q, lane) runs the M-point inverse over the inputs k1 = SPLIT k' + q of
            // column n2 = r + INC j, multiplies by e^{+2 pi i n' q / A} = conj(tw[n' q B]) and parks G_q[n'] in the H area
            const int grp = tid / T;
            if (grp < KEEP * SPLIT && active) {
                const int j = grp / SPLIT, qq = grp - j * SPLIT;
                const int n2 = r + INC * j;
                constexpr int BP = B | 1;
                float2 x[M];
                MVSIM_UNROLL
                for (int k = 0; k < M; ++k) x[k] = sm[lane + ((SPLIT * k + qq) * BP + n2) * T];
                RegSel<M, 1, kPackedStrided>::run(x);
                float2* g = smh + G0 + (j * SPLIT + qq) * M * T + lane;
                MVSIM_UNROLL
                for (int n = 0; n < M; ++n) g[n * T] = (n == 0 || SPLIT == 1) ? x[n] : cmulc(x[n], q.tw[n * qq * B]);
            }
        } else if (PH == 6) {
            // part 2: thread (column j, m, half h, lane) combines x[n' + M m] = sum_q G_q[n'] e^{+2 pi i m q / SPLIT} for its n' and
            // stores the kept planes (o = n2 + n1 B - crop0 = INC kz exactly)
            const int grp = tid / T;
            if (grp < KEEP * SPLIT * RS && active) {
                const int h = grp % RS, jm = grp / RS;
                const int j = jm / SPLIT, m = jm - j * SPLIT;
                const int n2 = r + INC * j;
                float2 c[SPLIT];
                MVSIM_UNROLL
                for (int qq = 1; qq < SPLIT; ++qq) c[qq] = q.tw[((m * qq) % SPLIT) * (S::N / SPLIT)];      // conj applied below
                const float2* g = smh + G0 + j * SPLIT * M * T + lane;
                float2* dst = q.u + tile * q.u_tstride + outer * q.ostride + lane;
                const unsigned e = (unsigned)q.estride32;
                constexpr int CNT = M / RS;
                MVSIM_UNROLL
                for (int i = 0; i < CNT; ++i) {
                    const int n = h * CNT + i;
                    float2 acc = g[n * T];
                    MVSIM_UNROLL
                    for (int qq = 1; qq < SPLIT; ++qq) {
                        const float2 v = cmulc(g[(qq * M + n) * T], c[qq]);
                        acc.x += v.x; acc.y += v.y;
                    }
                    const int o = n2 + (n + M * m) * B - q.crop0;
                    if ((unsigned)o < (unsigned)q.n_src) *at32(dst, umulhi32((uint32_t)o, q.keep_magic), e) = acc;
                }
            }
        } else {
            // the plane with the sum of the dropped slices: the A partial dot products with D' (phase 3), two chains
            if (p == 0 && active) {
                float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
                for (int k = 0; k + 1 < A; k += 2) {
                    s0.x += smh[k * T + lane].x; s0.y += smh[k * T + lane].y;
                    s1.x += smh[(k + 1) * T + lane].x; s1.y += smh[(k + 1) * T + lane].y;
                }
                if (A % 2) { s0.x += smh[(A - 1) * T + lane].x; s0.y += smh[(A - 1) * T + lane].y; }
                q.u[tile * q.u_tstride + outer * q.ostride + lane + q.n_keep * q.estride] = make_float2(s0.x + s1.x, s0.y + s1.y);
            }
        }
    }
};

// ==============================================================================================
// x passes.  A real row of padded length 2N is transformed with ONE N-point complex FFT:
//   z[m] = (r[m] - i r[m+N]) * exp(-i pi m / 2N)      (fold + twist)
// gives the even bins of the odd-frequency DFT  X[k] = sum_n r[n] exp(-2 pi i (k+1/2) n / 2N),
// which diagonalises negacyclic convolution.  The padded length is >= dim + kdim - 1, so no
// wrapped term reaches the cropped output and the result equals the reference's cyclic
// FFTConvolution inside the original interval -- with exactly N complex bins per row (no Nyquist
// column, no real-to-complex post-processing pass).
// ==============================================================================================
struct XParams {
    const float* rin;       // forward: real rows [n_rows][X]
    float* rout;            // inverse: real rows [n_rows][X]
    const float2* cin;      // inverse: spectrum rows [n_rows][N]
    float2* cout;           // forward: spectrum rows [n_rows][N]
    const float2* tw;       // exp(-2 pi i m / N)
    const float2* twist;    // exp(-i pi m / 2N)
    double* partials;       // inverse: per-block sums of the stored voxels of the rows >= sum_row0 (may be null)
    long long sum_row0;
    int X, n_rows, left, ext, crop0;
    int prefetch_dist;      // forward: thread 0 of CTA i prefetches the rows of CTA i + prefetch_dist into the L2 (0 = off)
    unsigned esz;           // sizeof(float) as a RUN-TIME value: element addresses of the gather become one widening multiply-add
                            // (IMAD.WIDE.U32) instead of the shift + add with carry pair a constant scale compiles to (144 of the
                            // forward pass's 450 gather instructions per warp were LEA, ncu r02j)
};

// element i of a real row through the run-time element size (see XParams::esz)
MVSIM_HD const float* rowat(const float* base, int i, unsigned esz)
{
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned long long)(unsigned)i * (unsigned long long)esz);
}

template <int A> struct XFwdState { float a[A], b[A]; };

// Forward x pass in three phases: (0) every thread gathers its 2 A samples into registers (indices first, then all loads),
// (1) fold + twist + first half, (2) second half + stores.  Splitting the gather from the transform at a barrier took the pass
// from 0.952 to 0.894 ms at config 3 (8 instead of 64 bytes of spills under the 80-register cap).  Measured and dropped in round 2:
// staging the CTA's ten contiguous rows in shared memory with one TMA bulk copy (0.903 ms: the L1 / LSU pressure of the 48 scalar
// loads per thread, 82 % of that pipe's peak under ncu, is not what bounds the pass).
template <int A_, int B_, int R_> struct XFwd : LineShape<A_, B_> {
    using S = LineShape<A_, B_>;
    static constexpr bool IS_X = true;
    static constexpr int A = A_, B = B_, R = R_;
    static constexpr int THREADS = R * S::P;
    static constexpr int SMEM_BYTES = S::ELEMS * R * (int)sizeof(float2);
    static constexpr int NPH = 3;
    using Params = XParams;
    using State = XFwdState<A_>;

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int, int tid, float2* sm, State& st)
    {
        constexpr int N = S::N;
        const int r = tid / S::P, p = tid % S::P;
        const long long row = (long long)bx * R + r;
        const bool active = row < q.n_rows;
        if (PH == 0) {
#ifdef __CUDA_ARCH__
            if (tid == 0 && q.prefetch_dist > 0) {
                const long long r0 = ((long long)bx + q.prefetch_dist) * R;
                if (r0 < q.n_rows) {
                    const long long nr = q.n_rows - r0 < R ? q.n_rows - r0 : R;
                    bulk_prefetch_l2(q.rin + r0 * q.X, (unsigned)(nr * q.X * sizeof(float)));
                }
            }
#endif
            if (p < B && active) {
                const float* __restrict__ src = q.rin + row * q.X;
                // indices first (branch free in the common modes), then all loads
                if (q.ext == EXT_MIRROR1) {
                    int ia[A], ib[A];
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        ia[n1] = mirror_once(p + n1 * B - q.left, q.X);
                        ib[n1] = mirror_once(p + n1 * B + N - q.left, q.X);
                    }
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) { st.a[n1] = *rowat(src, ia[n1], q.esz); st.b[n1] = *rowat(src, ib[n1], q.esz); }
                } else if (q.ext == EXT_ZERO) {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        const int i0 = p + n1 * B - q.left, i1 = i0 + N;
                        const bool ok0 = (unsigned)i0 < (unsigned)q.X, ok1 = (unsigned)i1 < (unsigned)q.X;
                        const float v0 = src[ok0 ? i0 : 0], v1 = src[ok1 ? i1 : 0];
                        st.a[n1] = ok0 ? v0 : 0.f;
                        st.b[n1] = ok1 ? v1 : 0.f;
                    }
                } else {
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        st.a[n1] = src[mirror_single(p + n1 * B - q.left, q.X)];
                        st.b[n1] = src[mirror_single(p + n1 * B + N - q.left, q.X)];
                    }
                }
            }
        } else if (PH == 1) {
            if (p < B && active) {
                float2 x[A];
                MVSIM_UNROLL
                for (int n1 = 0; n1 < A; ++n1) {
                    const float2 t = q.twist[p + n1 * B];
                    x[n1] = make_float2(st.a[n1] * t.x + st.b[n1] * t.y, st.a[n1] * t.y - st.b[n1] * t.x);   // (a - i b) * t
                }
                fwd_first<A, B, kPackedX>(p, x, sm, r * S::ELEMS, 1, q.tw);
            }
        } else {
            if (p < A && active) {
                float2 y[B];
                fwd_second<A, B, kPackedX>(p, y, sm, r * S::ELEMS, 1);
                float2* dst = q.cout + row * N;
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) dst[p + A * k2] = y[k2];
            }
        }
    }
};

template <int A_, int B_, int R_> struct XInv : LineShape<A_, B_> {
    using S = LineShape<A_, B_>;
    static constexpr bool IS_X = true;
    static constexpr int A = A_, B = B_, R = R_;
    static constexpr int THREADS = R * S::P;
    // exchange area + per-thread float partials + 32 double partials
    static constexpr int EXCH = S::ELEMS * R;
    static constexpr int SMEM_BYTES = EXCH * (int)sizeof(float2) + THREADS * (int)sizeof(float) + 32 * (int)sizeof(double) + 8;
    static constexpr int NPH = 4;
    using Params = XParams;
    using State = NoState;

    static MVSIM_HD float* psum(float2* sm) { return reinterpret_cast<float*>(sm + EXCH); }
    static MVSIM_HD double* dsum(float2* sm)
    {
        // 8-byte aligned: EXCH float2 (8 B each) then THREADS floats rounded up to an even count
        return reinterpret_cast<double*>(sm + EXCH + (THREADS + 1) / 2);
    }

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int, int tid, float2* sm, State&)
    {
        constexpr int N = S::N;
        const int r = tid / S::P, p = tid % S::P;
        const long long row = (long long)bx * R + r;
        const bool active = row < q.n_rows;
        if (PH == 0) {
#ifdef __CUDA_ARCH__
            if (tid == 0 && q.prefetch_dist > 0) {
                const long long r0 = ((long long)bx + q.prefetch_dist) * R;
                if (r0 < q.n_rows) {
                    const long long nr = q.n_rows - r0 < R ? q.n_rows - r0 : R;
                    bulk_prefetch_l2(q.cin + r0 * N, (unsigned)(nr * N * sizeof(float2)));
                }
            }
#endif
            if (p < A && active) {
                float2 y[B];
                const float2* src = q.cin + row * N;
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) y[k2] = src[p + A * k2];
                inv_first<A, B, kPackedX>(p, y, sm, r * S::ELEMS, 1, q.tw);
            }
        } else if (PH == 1) {
            float acc = 0.f;
            if (p < B && active) {
                float2 x[A];
                inv_second<A, B, kPackedX>(p, x, sm, r * S::ELEMS, 1);
                // one base pointer per thread (output index of n1 = 0), compile-time offsets n1 B and N + n1 B from it: the stores
                // carry immediate offsets instead of one address computation each
                float* d1 = q.rout + (row * q.X + (p - q.crop0));
                MVSIM_UNROLL
                for (int n1 = 0; n1 < A; ++n1) {
                    const int m = p + n1 * B;
                    const float2 t = q.twist[m];
                    // x * conj(t): real part -> padded m, minus imaginary part -> padded m + N
                    const float re = x[n1].x * t.x + x[n1].y * t.y;
                    const float mi = x[n1].x * t.y - x[n1].y * t.x;
                    const int o1 = m - q.crop0, o2 = m + N - q.crop0;
                    if ((unsigned)o1 < (unsigned)q.X) { d1[n1 * B] = re; acc += re; }
                    if ((unsigned)o2 < (unsigned)q.X) { d1[N + n1 * B] = mi; acc += mi; }
                }
            }
            if (q.partials) psum(sm)[tid] = row >= q.sum_row0 ? acc : 0.f;
        } else if (PH == 2) {
            if (q.partials && tid < 32) {
                double s = 0.0;
                for (int j = tid; j < THREADS; j += 32) s += (double)psum(sm)[j];
                dsum(sm)[tid] = s;
            }
        } else {
            if (q.partials && tid == 0) {
                double s = 0.0;
                for (int j = 0; j < 32; ++j) s += dsum(sm)[j];
                q.partials[bx] = s;
            }
        }
    }
};

}  // namespace mvsim

#include "zfused_poly.cuh"         // ZFusedPoly: the whole-view fused z pass in polyphase form
