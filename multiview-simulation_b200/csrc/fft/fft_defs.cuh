// Shared definitions for the line-FFT code.  Everything arithmetic is __host__ __device__ so the
// CPU emulation harness (tests/emu) executes exactly the code the kernels run.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define MVSIM_HD __host__ __device__ __forceinline__
#define MVSIM_UNROLL _Pragma("unroll")
#define MVSIM_UNROLL4 _Pragma("unroll 4")
#else
#define MVSIM_HD inline
#define MVSIM_UNROLL
#define MVSIM_UNROLL4
#ifndef MVSIM_HOST_FLOAT2
#define MVSIM_HOST_FLOAT2
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif
#endif

namespace mvsim {

// Blackwell's packed FP32x2 arithmetic (crt/sm_100_rt.h: FADD2 / FMUL2 / FFMA2 in SASS); plain float pairs in the CPU emulation
MVSIM_HD float2 add2(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
MVSIM_HD float2 mul2(float2 a, float2 b)
{
#ifdef __CUDA_ARCH__
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
MVSIM_HD float2 fma2(float2 a, float2 b, float2 c)
{
#ifdef __CUDA_ARCH__
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

MVSIM_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
MVSIM_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

// Views.extendMirrorSingle index map (imglib2): period 2(n-1), border sample not repeated; n==1 -> 0.
MVSIM_HD int mirror_single(int i, int n)
{
    if ((unsigned)i < (unsigned)n) return i;
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    int j = i % p;
    if (j < 0) j += p;
    return j >= n ? p - j : j;
}

MVSIM_HD uint32_t umulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
// floor(n / d) == umulhi32(n, div_magic(d)) for n, d < 2^16
inline uint32_t div_magic(uint32_t d) { return (uint32_t)((0x100000000ull + d - 1) / d); }

// Extension modes of the line loaders.  EXT_MIRROR1 is the common case (every padded index folds at most
// once): branch free, so the unrolled loads of a thread are issued back to back.
enum { EXT_MIRROR1 = 0, EXT_ZERO = 1, EXT_MIRROR_GENERAL = 2 };

// valid when -(n-1) <= i <= 2(n-1)
MVSIM_HD int mirror_once(int i, int n)
{
    const int a = i < 0 ? -i : i;
    const int b = 2 * (n - 1) - a;
    return a < b ? a : b;
}

// host side: which mode may a loader use for padded length `npad`, margin `left`, source length n?
inline int mirror_mode(int npad, int left, int n)
{
    return (n > 1 && left <= n - 1 && npad - 1 - left <= 2 * (n - 1)) ? EXT_MIRROR1 : EXT_MIRROR_GENERAL;
}

// 16-byte asynchronous global->shared copy (LDGSTS); plain copy in the CPU emulation
MVSIM_HD void cp_async16(void* smem_dst, const void* gmem_src)
{
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
#else
    const float4* s4 = static_cast<const float4*>(gmem_src);
    *static_cast<float4*>(smem_dst) = *s4;
#endif
}
MVSIM_HD void cp_async_wait_all()
{
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
#endif
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier, sm_100a.  Device only; the CPU emulation copies directly. ----
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n.reg .pred p;\nMVSIM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MVSIM_DONE;\nbra MVSIM_WAIT;\nMVSIM_DONE:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// 4-D tiled tensor load global -> shared, completion signalled on the mbarrier
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// L2 prefetches issued by ONE thread for the data a later CTA will gather: a whole tile (tensor form) or a contiguous range
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];\n" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gptr), "r"(bytes) : "memory");
}
#endif

}  // namespace mvsim
