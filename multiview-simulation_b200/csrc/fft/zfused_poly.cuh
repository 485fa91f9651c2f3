// Fused z pass of the whole-view call in POLYPHASE form (included at the end of line_fft.cuh).
//
// extractSlices keeps the planes z = 0, INC, 2 INC, ... of the convolved volume (S/SimulateMultiViewDataset.java:206) and
// adjustImage (S/Tools.java:143-147) needs, of all the other planes, only their sum.  With the padded z line length N = INC * M
// the kept outputs of the cyclic convolution c = a' (*) b  (a' = mirror-extended image line, b = zero-extended PSF line) are
//
//     c[crop0 + INC k] = sum_s ( beta_s (*)_M alpha_s )[k],     alpha_s[m] = a'[(crop0 - s + INC m) mod N],  beta_s[q] = b[INC q + s],
//
// i.e. INC cyclic convolutions of length M = N / INC.  So the pass runs INC M-point forward transforms of the image phases, INC
// (heavily pruned: ceil(KZ / INC) of M samples are non-zero) M-point transforms of the PSF phases, INC M multiply-adds, a sum over
// the phases and ONE M-point inverse -- no radix-INC combine stage on either forward side and a fifth (third) of an inverse,
// against three N-point transforms of the spectral kernels (ZFusedOTF) or two and a pruned one (ZFusedDec).
//
// The extra plane -- the sum of ALL cropped planes, see the end of this comment -- comes from the time domain: with
// Bpre[i] = sum_{j < i} b[j], B0 = Bpre[KZ],
//
//     sum_{o < n_src} c[crop0 + o] = B0 * sum_{n < n_src} a'[n] + sum_{i = 1}^{KZ-1} Bpre[i] (a'[n_src + KZ-1 - i] - a'[KZ-1 - i])
//
// (each a'[n] is weighted by the part of the PSF line whose taps land inside the cropped range), and
// sum_{n < n_src} a'[n] = sum_s DFT(alpha_s)[0] - sum_{n >= n_src} a'[n].  The 2 (KZ - 1) border samples are loaded a second time
// (L1 / L2 hits: the same CTA gathers them in the same phase): group A + c owns the taps j = 4 c .. 4 c + 3, keeps their samples in
// registers, and in the next phase (PSF tile present) forms the differences DD[j], its local share
// L_c = sum_u (b[4c] + .. + b[4c+u-1]) DD[4c+u] and its chunk sums CS_c = sum b, DS_c = sum DD.  What is left,
// sum_c (CS_0 + .. + CS_{c-1}) DS_c + L_c, is a scan over the NCH chunks, run by ONE warp beside the two warps of the first
// inverse half.
//
// Work items of a CTA (T neighbouring kx columns of one ky; G = THREADS / T groups, group g works on `lane` = tid % T):
//   level 1 : (phase s, n2 < B): A-point transform over n1 of alpha_s[n1 B + n2]      -> INC B items, R1 rounds over the G groups
//   level 2 : (phase s, k1 < A): B-point transform                                     -> INC A items <= G, one per group
// so every warp works in the four transform phases; only the single inverse runs on A, then B groups.
// Phases: requests (PSF tile by TMA, border samples, gather) | level 1 of the PSF and image phases (shared twiddles), chunk terms |
//         level 2 of both, multiply, inverse level 1 of the item's OWN products | inverse level 2 (summing the INC phases as it
//         loads) + stores, chunk scan | sum plane.
// The inverse is linear, so the sum over the phases may come after its first half: every level-2 thread inverts the products it
// holds in registers (INC times the work of inverting the sum, but on all warps and without the two barriers and the two
// low-occupancy phases -- reduce, 2-warp inverse -- that cost 16 % of the kernel's warp samples in ncu r02g; the run time is
// the same, 1.81 ms, the tail simpler).
// The extra plane carries the sum of ALL cropped planes (conv_middle_z reports it): the inverse x pass then takes adjustImage's mean
// from that plane alone (XParams::sum_row0), no bookkeeping of the kept planes here.
#pragma once

namespace mvsim {

// groups per CTA for the split M = a * b (a <= b) with inc phases
constexpr int poly_groups(int a, int b, int inc)
{
    // one level-1 item per group up to 40 groups (320 threads at T = 8: two CTAs per SM at 96 registers), else two rounds
    const int l1 = inc * b, l2 = inc * a;
    if (l1 <= 40) return l1;
    const int half = (l1 + 1) / 2;
    return l2 > half ? l2 : half;
}
constexpr int kPolyTaps = 4;        // PSF taps per border group (their border samples live in registers between two phases)
// may a z line of n points run ZFusedPoly<n, inc, T>?  (inc phases of a supported length n / inc, enough groups for the border terms)
constexpr bool zfused_poly_ok(int n, int inc)
{
    if (!(inc == 3 || inc == 5) || n % inc != 0) return false;
    const FftSize s = fft_size_lookup(n / inc);
    if (s.n < 64 || s.a > s.b) return false;
    const int g = poly_groups(s.a, s.b, inc);
    // (two gathered items of more than 8 samples each do not fit the 96 registers of 2 x 320 threads per SM: such splits spill.
    // Measured with the spills at 576 = 3 x 192 points, inc 3: 0.514 ms against 0.502 ms for ZFusedDec<24,24,8,3> -- excluded)
    const int r1 = (inc * s.b + g - 1) / g;
    return g >= inc * s.a && g >= s.b + 4 && g - s.a >= 8 && g <= 40 && (r1 * s.a <= 16 || g <= 32);
}
// PSF taps the border groups of a launch can take
constexpr int zfused_poly_max_taps(int n, int inc)
{
    const FftSize s = fft_size_lookup(n / inc);
    return kPolyTaps * (poly_groups(s.a, s.b, inc) - s.a);
}

// shared memory without the PSF tile (the host needs it without the template)
constexpr int poly_smem_base(int a, int b, int inc, int t)
{
    const int g = poly_groups(a, b, inc), aux_rows = 4 * (g - a) + inc + 8 + 1;
    return ((2 * inc * a * (b | 1) * t + aux_rows * t + 15) / 16 * 16 + 16) * (int)sizeof(float2);
}
constexpr int poly_psf_tile_bytes(int k_src, int t) { return (k_src + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows * t * (int)sizeof(float2); }
constexpr int kPolySmemMax = (kSmemLimit - 2048) / 2;      // two CTAs per SM (1 KB reserved per CTA)
// the PSF tile (TMA) fits beside the exchange areas without costing the second resident CTA, and the border groups cover the taps
inline bool zfused_poly_fits(int n, int inc, int t, int k_src)
{
    const FftSize m = fft_size_lookup(n / inc);
    return k_src <= zfused_poly_max_taps(n, inc) && poly_smem_base(m.a, m.b, inc, t) + poly_psf_tile_bytes(k_src, t) <= kPolySmemMax;
}

template <int A, int B, int R1> struct PolyState {
    float2 x[R1][A];                // gathered image samples of the group's level-1 items (requested in phase 0, used in phase 1)
    float2 tl[kPolyTaps], hd[kPolyTaps];    // border samples of the group's taps
    float2 y[B];                    // image spectrum A_s[k1 + A k2] of the group's level-2 item
};

template <int N_, int INC_, int T_> struct ZFusedPoly {
    static constexpr bool IS_X = false;
    static constexpr int N = N_, INC = INC_, T = T_, M = N_ / INC_;
    static constexpr FftSize SZ = fft_size_lookup(M);
    static constexpr int A = SZ.a, B = SZ.b, BP = B | 1;
    static constexpr int L1 = INC * B, L2 = INC * A;
    static constexpr int G = poly_groups(A, B, INC);
    static constexpr int R1 = (L1 + G - 1) / G;           // level-1 items per group
    static constexpr int THREADS = G * T;
    static constexpr int MIN_BLOCKS = 2;
    static constexpr int NCH = G - A;                     // border groups: A .. G-1 (they idle during the first inverse half)
    static constexpr int NSC = 4, QSC = (NCH + NSC - 1) / NSC;      // the chunk scan runs on NSC groups (one warp at T = 8)
    static_assert(N_ % INC_ == 0 && A * B == M && G >= L2 && G >= B + NSC && NCH >= NSC, "unsupported split");
    // shared memory (float2 elements): [E1: INC x (A x BP) x T][E2: the same][AUX rows of T][mbarrier][PSF tile (TMA)]
    static constexpr int E_ELEMS = INC * A * BP * T;
    static constexpr int AUX_CS = 0, AUX_DS = NCH, AUX_L = 2 * NCH, AUX_TS = 3 * NCH, AUX_TOT = 4 * NCH, AUX_SC = AUX_TOT + INC,
                         AUX_B0 = AUX_SC + 2 * NSC, AUX_ROWS = AUX_B0 + 1;
    static constexpr int AUX0 = 2 * E_ELEMS;
    static constexpr int BAR_ELEMS = (AUX0 + AUX_ROWS * T + 15) / 16 * 16;
    static constexpr int PSF_ELEMS0 = BAR_ELEMS + 16;
    static constexpr int SMEM_BYTES = PSF_ELEMS0 * (int)sizeof(float2);       // without the PSF tile
    static constexpr int NPH = 5;
    using Params = ZFusedParams;
    using State = PolyState<A, B, R1>;
    static_assert(SMEM_BYTES == poly_smem_base(A, B, INC_, T_), "host-side shared memory formula out of sync");
    static int smem_bytes(const Params& q) { return SMEM_BYTES + (q.use_tma ? poly_psf_tile_bytes(q.k_src, T) : 0); }
    static int smem_bytes_max() { return kPolySmemMax > SMEM_BYTES ? kPolySmemMax : SMEM_BYTES; }

    // PSF sample j of this thread's column: TMA-fetched tile in shared memory (rows beyond KZ are zero filled by the unit); global in
    // the CPU emulation
    static MVSIM_HD float2 psf_at(const Params& q, const float2* sm, const float2* __restrict__ gsrc, int lane, int j)
    {
#ifdef __CUDA_ARCH__
        // on the device the kernel is only launched with the TMA-fed tile (the host falls back to the spectral kernels otherwise)
        (void)gsrc;
        const int rows = (q.k_src + kTmaBoxRows - 1) / kTmaBoxRows * kTmaBoxRows;
        return j < rows ? sm[PSF_ELEMS0 + j * T + lane] : make_float2(0.f, 0.f);
#else
        (void)sm; (void)lane;
        const bool ok = j < q.k_src;
        const float2 v = gsrc[(long long)(ok ? j : 0) * q.estride];
        return ok ? v : make_float2(0.f, 0.f);
#endif
    }

    template <int PH> static MVSIM_HD void phase(const Params& q, int bx, int by, int tid, float2* sm, State& st)
    {
        const int lane = tid % T, g = tid / T;
        const int tile = by, outer = bx;
        const bool active = (tile + q.tile0) * T + lane < q.kx_count;
        float2* e1 = sm;
        float2* e2 = sm + E_ELEMS;
        float2* aux = sm + AUX0 + lane;            // row r of the AUX area: aux[r * T]
        const float2* __restrict__ src = q.u + tile * q.u_tstride + outer * q.ostride + lane;
        const float2* __restrict__ psrc = q.p2 + tile * q.p2_tstride + outer * q.ostride + lane;
        if (PH == 0) {
            // requests only: PSF tile (TMA), L2 prefetch of a later CTA's lines, border samples, the gather.  Nothing is consumed
            // before the barrier, which also publishes the mbarrier's initialisation
#ifdef __CUDA_ARCH__
            if (tid == 0) {
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
            if (active) {
                const unsigned e = (unsigned)q.estride32;
                // border samples of the taps j = 4 c + u of group A + c (the host guarantees EXT_MIRROR1 and KZ <= 4 NCH):
                //   tail a'[top - j] (source top - j - left >= 0, folds at most once at the far end), head a'[KZ-1 - j] (source |..|)
                // (clamped index + select)
                if (g >= A) {
                    const int c = g - A, top = q.n_src + q.k_src - 1;
                    MVSIM_UNROLL
                    for (int u = 0; u < kPolyTaps; ++u) {
                        const int j = kPolyTaps * c + u;
                        const int it = top - j - q.left, ih = q.k_src - 1 - j - q.left;
                        const int mt = 2 * (q.n_src - 1) - it;
                        const bool okt = j < q.k_src && top - j < N, okh = j < q.k_src && j >= 1;
                        st.tl[u] = *at32(src, (unsigned)(okt ? (it < mt ? it : mt) : 0), e);
                        st.hd[u] = *at32(src, (unsigned)(okh ? (ih < 0 ? -ih : ih) : 0), e);
                    }
                }
                MVSIM_UNROLL
                for (int r = 0; r < R1; ++r) {
                    const int item = g + G * r;
                    if (R1 * G == L1 || item < L1) {
                        const int s = item / B, n2 = item - s * B;
                        int b0 = q.crop0 - s + INC * n2;          // alpha_s[n1 B + n2] = a'[(b0 + INC B n1) mod N]
                        b0 = b0 < 0 ? b0 + N : b0;
                        b0 = b0 >= N ? b0 - N : b0;
                        int idx[A];
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < A; ++n1) {
                            const int nn = b0 + INC * B * n1;
                            idx[n1] = mirror_once((nn >= N ? nn - N : nn) - q.left, q.n_src);
                        }
                        MVSIM_UNROLL
                        for (int n1 = 0; n1 < A; ++n1) st.x[r][n1] = *at32(src, (unsigned)idx[n1], e);
                    }
                }
            }
        } else if (PH == 1) {
#ifdef __CUDA_ARCH__
            mbar_wait(reinterpret_cast<uint64_t*>(sm + BAR_ELEMS), 0);       // the PSF tile (8 KB) lands before the gathered lines
#endif
            if (active) {
                // level 1 of the PSF phases (beta_s[q] = INC b[INC q + s], non-zero for q < ceil(KZ / INC) only) and of the image
                // phases, item by item: both use the twiddles W_M^{k1 n2} of the item
                constexpr int K = RegSelZ<A, kPackedStrided>::K;
                const bool pruned = (q.k_src + INC - 1) / INC <= K * B;
                MVSIM_UNROLL
                for (int r = 0; r < R1; ++r) {
                    const int item = g + G * r;
                    if (R1 * G == L1 || item < L1) {
                        const int s = item / B, n2 = item - s * B;
                        float2 tw[A];
                        MVSIM_UNROLL
                        for (int k1 = 1; k1 < A; ++k1) tw[k1] = q.tw[INC * k1 * n2];
                        float2 x[A];
                        if (pruned) {
                            MVSIM_UNROLL
                            for (int n1 = 0; n1 < K; ++n1) {
                                const float2 v = psf_at(q, sm, psrc, lane, INC * (n1 * B + n2) + s);
                                x[n1] = make_float2(v.x * (float)INC, v.y * (float)INC);
                            }
                            RegSelZ<A, kPackedStrided>::run(x);
                        } else {
                            MVSIM_UNROLL
                            for (int n1 = 0; n1 < A; ++n1) {
                                const float2 v = psf_at(q, sm, psrc, lane, INC * (n1 * B + n2) + s);
                                x[n1] = make_float2(v.x * (float)INC, v.y * (float)INC);
                            }
                            RegSel<A, -1, kPackedStrided>::run(x);
                        }
                        float2* prow = e2 + ((s * A) * BP + n2) * T + lane;
                        MVSIM_UNROLL
                        for (int k1 = 0; k1 < A; ++k1) prow[k1 * BP * T] = k1 == 0 ? x[0] : cmul(x[k1], tw[k1]);
                        RegSel<A, -1, kPackedStrided>::run(st.x[r]);
                        float2* row = e1 + ((s * A) * BP + n2) * T + lane;
                        MVSIM_UNROLL
                        for (int k1 = 0; k1 < A; ++k1) row[k1 * BP * T] = k1 == 0 ? st.x[r][0] : cmul(st.x[r][k1], tw[k1]);
                    }
                }
                if (g >= A) {
                    // border terms of chunk c: DD[u] = tail - head (DD = 0 at tap 0 and beyond the PSF line), the chunk sums
                    // TS_c (tail samples), DS_c (differences), CS_c (PSF taps) and L_c = sum_u (taps of the chunk before u) DD[u]
                    const int c = g - A, top = q.n_src + q.k_src - 1;
                    float2 ts = make_float2(0.f, 0.f), ds = make_float2(0.f, 0.f), pre = make_float2(0.f, 0.f), loc = make_float2(0.f, 0.f);
                    MVSIM_UNROLL
                    for (int u = 0; u < kPolyTaps; ++u) {
                        const int j = kPolyTaps * c + u;
                        const bool okt = j < q.k_src && top - j < N, okh = j < q.k_src && j >= 1;
                        const float2 t = okt ? st.tl[u] : make_float2(0.f, 0.f);
                        const float2 d = okh ? make_float2(t.x - st.hd[u].x, t.y - st.hd[u].y) : make_float2(0.f, 0.f);
                        ts.x += t.x; ts.y += t.y;
                        ds.x += d.x; ds.y += d.y;
                        loc.x += pre.x * d.x - pre.y * d.y;
                        loc.y += pre.x * d.y + pre.y * d.x;
                        const float2 b = psf_at(q, sm, psrc, lane, j);
                        pre.x += b.x; pre.y += b.y;
                    }
                    // padding samples beyond the last tap's reach (the planner's size is rarely exactly n_src + KZ - 1)
                    for (int n = top + 1 + c; n < N; n += NCH) {
                        const float2 v = *at32(src, (unsigned)mirror_once(n - q.left, q.n_src), (unsigned)q.estride32);
                        ts.x += v.x; ts.y += v.y;
                    }
                    aux[(AUX_TS + c) * T] = ts;
                    aux[(AUX_DS + c) * T] = ds;
                    aux[(AUX_CS + c) * T] = pre;
                    aux[(AUX_L + c) * T] = loc;
                }
            }
        } else if (PH == 2) {
            // level 2 of the image phase (spectrum A_s[k1 + A k2] in registers; the DC bins give the sum of the whole padded line) and
            // of the PSF phase, multiply; then the first half of the inverse on the item's own products: B-point inverse over k2,
            // twiddle, into this item's row of E1 (read completely by this thread before it is rewritten)
            if (g < L2 && active) {
                float2* row = e1 + g * BP * T + lane;               // item g = s A + k1
                MVSIM_UNROLL
                for (int n2 = 0; n2 < B; ++n2) st.y[n2] = row[n2 * T];
                RegSel<B, -1, kPackedStrided>::run(st.y);
                if (g % A == 0) aux[(AUX_TOT + g / A) * T] = st.y[0];
                const float2* prow = e2 + g * BP * T + lane;
                float2 y[B];
                MVSIM_UNROLL
                for (int n2 = 0; n2 < B; ++n2) y[n2] = prow[n2 * T];
                RegSel<B, -1, kPackedStrided>::run(y);
                MVSIM_UNROLL
                for (int k2 = 0; k2 < B; ++k2) y[k2] = cmul(y[k2], st.y[k2]);
                RegSel<B, 1, kPackedStrided>::run(y);
                const int k1 = g % A;
                MVSIM_UNROLL
                // (the twiddle does not depend on the phase, so it could follow the sum in the next phase -- a quarter of the multiplies,
                // but on 4 warps instead of 10: measured 1.813 -> 1.809 ms, not worth the longer tail)
                for (int n2 = 0; n2 < B; ++n2) row[n2 * T] = n2 == 0 ? y[0] : cmulc(y[n2], q.tw[INC * n2 * k1]);
            }
        } else if (PH == 3) {
            if (g < B) {
                // inverse level 2: thread n2 sums the INC phases as it loads, A-point inverse, and ends up with
                // c[crop0 + INC (n2 + n1 B)]: the kept planes kz = n2 + n1 B, compacted
                if (active) {
                    float2 x[A];
                    const float2* col = e1 + g * T + lane;
                    MVSIM_UNROLL
                    for (int k1 = 0; k1 < A; ++k1) {
                        float2 acc = col[k1 * BP * T];
                        MVSIM_UNROLL
                        for (int s = 1; s < INC; ++s) { const float2 v = col[(s * A + k1) * BP * T]; acc.x += v.x; acc.y += v.y; }
                        x[k1] = acc;
                    }
                    RegSel<A, 1, kPackedStrided>::run(x);
                    float2* dst = q.u + tile * q.u_tstride + outer * q.ostride + lane;
                    const unsigned e = (unsigned)q.estride32;
                    MVSIM_UNROLL
                    for (int n1 = 0; n1 < A; ++n1) {
                        const int kz = g + n1 * B;
                        if (kz < q.n_keep) *at32(dst, (unsigned)kz, e) = x[n1];
                    }
                }
            } else if (g >= G - NSC && active) {
                // chunk scan, quarter i: sum_c (CS_0 + .. + CS_{c-1}) DS_c + L_c and the tail sums over the chunks [i QSC, (i + 1) QSC)
                const int i = g - (G - NSC), c0 = i * QSC, c1 = (i + 1) * QSC < NCH ? (i + 1) * QSC : NCH;
                float2 pre = make_float2(0.f, 0.f);
                for (int k = 0; k < c0; ++k) { const float2 v = aux[(AUX_CS + k) * T]; pre.x += v.x; pre.y += v.y; }
                float2 dot = make_float2(0.f, 0.f), ts = make_float2(0.f, 0.f);
                for (int c = c0; c < c1; ++c) {
                    const float2 d = aux[(AUX_DS + c) * T], l = aux[(AUX_L + c) * T], t = aux[(AUX_TS + c) * T], b = aux[(AUX_CS + c) * T];
                    dot.x += pre.x * d.x - pre.y * d.y + l.x;
                    dot.y += pre.x * d.y + pre.y * d.x + l.y;
                    ts.x += t.x; ts.y += t.y;
                    pre.x += b.x; pre.y += b.y;
                }
                if (i == NSC - 1) aux[AUX_B0 * T] = pre;        // B0: the whole PSF line
                aux[(AUX_SC + i) * T] = dot;
                aux[(AUX_SC + NSC + i) * T] = ts;
            }
        } else {
            // plane n_keep: N (B0 sum_{n < n_src} a'[n] + dot), the sum of ALL cropped planes (the z transforms are unnormalised: factor N
            // as in the spectral kernels; the kept planes carry it through beta's factor INC and the unnormalised M-point inverse)
            if (g == 0 && active) {
                float2 d = make_float2(0.f, 0.f), t = make_float2(0.f, 0.f), tot = make_float2(0.f, 0.f);
                MVSIM_UNROLL
                for (int j = 0; j < NSC; ++j) {
                    const float2 a = aux[(AUX_SC + j) * T], b = aux[(AUX_SC + NSC + j) * T];
                    d.x += a.x; d.y += a.y; t.x += b.x; t.y += b.y;
                }
                MVSIM_UNROLL
                for (int s = 0; s < INC; ++s) { const float2 v = aux[(AUX_TOT + s) * T]; tot.x += v.x; tot.y += v.y; }
                const float2 b0 = aux[AUX_B0 * T];
                const float2 sa = make_float2(tot.x - t.x, tot.y - t.y);
                const float2 all = make_float2(b0.x * sa.x - b0.y * sa.y + d.x, b0.x * sa.y + b0.y * sa.x + d.y);
                q.u[tile * q.u_tstride + outer * q.ostride + lane + q.n_keep * q.estride] = make_float2((float)N * all.x, (float)N * all.y);
            }
        }
    }
};

}  // namespace mvsim
