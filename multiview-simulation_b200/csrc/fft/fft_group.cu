// Instantiates the line-FFT kernels (line_fft.cuh) for one group of sizes: -DMVSIM_GROUP=0..4.
#include <type_traits>

#include "fft_launch.h"

#if !defined(MVSIM_GROUP) || !defined(MVSIM_LANES)
#error "compile with -DMVSIM_GROUP=<0..4> -DMVSIM_LANES=<4|8>"
#endif

namespace mvsim {

// resident CTAs per SM the register allocator must allow, so that the load, exchange and store phases of
// different tiles overlap: x passes 4 (<= 192 threads: forward 144, inverse 192) / 3 (<= 256 threads), strided passes 2 (T=8) / 4 (T=4).
template <class K> constexpr int min_blocks()
{
    return K::IS_X ? (K::THREADS <= 192 ? 4 : (K::THREADS <= 256 ? 3 : 1)) : (K::THREADS <= 160 ? 4 : (K::THREADS <= 288 ? 2 : 1));
}

// a kernel may state its own residency target (K::MIN_BLOCKS)
template <class K, class = void> struct MinBlocksOf { static constexpr int value = min_blocks<K>(); };
template <class K> struct MinBlocksOf<K, std::void_t<decltype(K::MIN_BLOCKS)>> { static constexpr int value = K::MIN_BLOCKS; };

template <class K> __global__ void __launch_bounds__(K::THREADS, MinBlocksOf<K>::value) fft_kernel(const __grid_constant__ typename K::Params q)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* sm = reinterpret_cast<float2*>(smraw);
    typename K::State st;
    K::template phase<0>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st);
    if constexpr (K::NPH > 1) { __syncthreads(); K::template phase<1>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
    if constexpr (K::NPH > 2) { __syncthreads(); K::template phase<2>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
    if constexpr (K::NPH > 3) { __syncthreads(); K::template phase<3>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
    if constexpr (K::NPH > 4) { __syncthreads(); K::template phase<4>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
    if constexpr (K::NPH > 5) { __syncthreads(); K::template phase<5>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
    if constexpr (K::NPH > 6) { __syncthreads(); K::template phase<6>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
    if constexpr (K::NPH > 7) { __syncthreads(); K::template phase<7>(q, blockIdx.x, blockIdx.y, threadIdx.x, sm, st); }
}

// dynamic shared memory of a launch: K::SMEM_BYTES, or K::smem_bytes(params) where the kernel sizes an area at run time
template <class K, class = void> struct SmemOf {
    static int bytes(const typename K::Params&) { return K::SMEM_BYTES; }
    static int max_bytes() { return K::SMEM_BYTES; }
};
template <class K> struct SmemOf<K, std::void_t<decltype(&K::smem_bytes_max)>> {
    static int bytes(const typename K::Params& q) { return K::smem_bytes(q); }
    static int max_bytes() { return K::smem_bytes_max(); }
};

template <class K> static int launch(const void* params, unsigned gx, unsigned gy, cudaStream_t s)
{
    static bool configured = false;     // per instantiation; idempotent, so a race only repeats the call
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fft_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemOf<K>::max_bytes());
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const typename K::Params& q = *static_cast<const typename K::Params*>(params);
    fft_kernel<K><<<dim3(gx, gy), K::THREADS, SmemOf<K>::bytes(q), s>>>(q);
    return (int)cudaGetLastError();
}

#ifndef MVSIM_DEC_UNIT
#define MVSIM_DEC_UNIT 0
#endif
#ifndef MVSIM_POLY_UNIT
#define MVSIM_POLY_UNIT 0
#endif

template <int A, int B> static int launch_size(int kind, const void* params, unsigned gx, unsigned gy, cudaStream_t s)
{
    constexpr int RF = x_rows_per_block(A, B, false), RI = x_rows_per_block(A, B, true);
    constexpr int T = MVSIM_LANES;
    (void)RF; (void)RI;
#if MVSIM_POLY_UNIT
    constexpr bool built = A * B >= kDecMinLine && A * B <= kDecMaxLine;
    switch (kind) {
    case FFT_ZFUSED_POLY3:
        if constexpr (built && zfused_poly_ok(A * B, 3)) return launch<ZFusedPoly<A * B, 3, T>>(params, gx, gy, s);
        break;
    case FFT_ZFUSED_POLY5:
        if constexpr (built && zfused_poly_ok(A * B, 5)) return launch<ZFusedPoly<A * B, 5, T>>(params, gx, gy, s);
        break;
    }
#elif MVSIM_DEC_UNIT
    // the decimated fused z kernels live in their own translation units (build time)
    constexpr bool built = A * B >= kDecMinLine && A * B <= kDecMaxLine;
    switch (kind) {
    case FFT_ZFUSED_DEC3:
        if constexpr (built && zfused_dec_ok(A, B, 3)) return launch<ZFusedDec<B, A, T, 3>>(params, gx, gy, s);
        break;
    case FFT_ZFUSED_DEC5:
        if constexpr (built && zfused_dec_ok(A, B, 5)) return launch<ZFusedDec<B, A, T, 5>>(params, gx, gy, s);
        break;
    }
#else
    switch (kind) {
#if MVSIM_LANES == 8
    case FFT_XFWD: return launch<XFwd<A, B, RF>>(params, gx, gy, s);
    case FFT_XINV: return launch<XInv<A, B, RI>>(params, gx, gy, s);
#endif
    case FFT_SFWD: return launch<StridedFwd<A, B, T>>(params, gx, gy, s);
    case FFT_SINV: return launch<StridedInv<A, B, T>>(params, gx, gy, s);
    case FFT_ZFUSED: return launch<ZFused<A, B, T>>(params, gx, gy, s);
    case FFT_ZFUSED_OTF: return launch<ZFusedOTF<A, B, T>>(params, gx, gy, s);
    }
#endif
    return (int)cudaErrorInvalidValue;
}

#define MVSIM_CAT_(a, b) a##b
#define MVSIM_CAT(a, b) MVSIM_CAT_(a, b)

#if MVSIM_GROUP == 0
#define MVSIM_GROUP_SIZES(X) MVSIM_FFT_SIZES_SMALL(X)
#elif MVSIM_GROUP == 1
#define MVSIM_GROUP_SIZES(X) MVSIM_FFT_SIZES_G1(X)
#elif MVSIM_GROUP == 2
#define MVSIM_GROUP_SIZES(X) MVSIM_FFT_SIZES_G2(X)
#elif MVSIM_GROUP == 3
#define MVSIM_GROUP_SIZES(X) MVSIM_FFT_SIZES_G3(X)
#else
#define MVSIM_GROUP_SIZES(X) MVSIM_FFT_SIZES_G4(X)
#endif

#if MVSIM_POLY_UNIT
#define MVSIM_FN MVSIM_CAT(MVSIM_CAT(MVSIM_CAT(fft_launch_poly_g, MVSIM_GROUP), _t), MVSIM_LANES)
#elif MVSIM_DEC_UNIT
#define MVSIM_FN MVSIM_CAT(MVSIM_CAT(MVSIM_CAT(fft_launch_dec_g, MVSIM_GROUP), _t), MVSIM_LANES)
#else
#define MVSIM_FN MVSIM_CAT(MVSIM_CAT(MVSIM_CAT(fft_launch_g, MVSIM_GROUP), _t), MVSIM_LANES)
#endif
int MVSIM_FN(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s)
{
    switch (n) {
#define MVSIM_X(n_, a_, b_) case n_: return launch_size<a_, b_>(kind, params, gx, gy, s);
        MVSIM_GROUP_SIZES(MVSIM_X)
#undef MVSIM_X
    }
    return -1;
}

}  // namespace mvsim
