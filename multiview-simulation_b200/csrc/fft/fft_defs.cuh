// Shared definitions for the line-FFT code.  Everything arithmetic is __host__ __device__ so the
// CPU emulation harness (tests/emu) executes exactly the code the kernels run.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define MVSIM_HD __host__ __device__ __forceinline__
#define MVSIM_UNROLL _Pragma("unroll")
#else
#define MVSIM_HD inline
#define MVSIM_UNROLL
#ifndef MVSIM_HOST_FLOAT2
#define MVSIM_HOST_FLOAT2
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif
#endif

namespace mvsim {

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

}  // namespace mvsim
