// Counter-based Poisson sampler: Philox4x32-10 + exact Poisson variates.
//
// Replaces Tools.poissonProcess / PoissonGenerator.nextValue (S/Tools.java:73-86,
// S/uncommons/PoissonGenerator.java:95-109): the reference counts exponential inter-arrival times
// drawn from ONE sequential java.util.Random (O(lambda) Math.log calls per voxel, inherently
// serial).  Here every output voxel owns a Philox counter (key = seed, counter = voxel index,
// stream id, attempt), so the result does not depend on the launch geometry or on how views are
// sharded over GPUs.  The variate is exact Poisson(lambda):
//   lambda < 10 : inversion by sequential search on one 53-bit uniform,
//   lambda >= 10: PTRS transformed rejection (W. Hoermann, Insurance: Mathematics and Economics 12, 1993).
// lambda <= 0 or NaN gives 0 (the reference loop does not terminate for lambda < 0).
// __host__ __device__: tests/emu evaluates the same code on the CPU.
#pragma once
#include "fft/fft_defs.cuh"

namespace mvsim {

struct Philox4 { uint32_t x, y, z, w; };

MVSIM_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MVSIM_HD Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1)
{
    MVSIM_UNROLL
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        Philox4 n;
        n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

// (0,1) uniform from 53 of 64 random bits
MVSIM_HD double u01(uint32_t hi, uint32_t lo)
{
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)v + 0.5) * 0x1.0p-53;
}

struct PoissonKey { uint32_t k0, k1, stream_lo; };

MVSIM_HD PoissonKey make_poisson_key(uint64_t seed, uint64_t stream)
{
    PoissonKey k;
    k.k0 = (uint32_t)seed;
    k.k1 = (uint32_t)(seed >> 32) ^ ((uint32_t)(stream >> 32) * 0x9E3779B9u);
    k.stream_lo = (uint32_t)stream;
    return k;
}

MVSIM_HD float poisson_sample(double lam, uint64_t index, PoissonKey key)
{
    if (!(lam > 0.0)) return 0.f;
    Philox4 c;
    c.x = (uint32_t)index; c.y = (uint32_t)(index >> 32); c.z = key.stream_lo; c.w = 0;
    if (lam < 10.0) {
        const Philox4 r = philox4x32_10(c, key.k0, key.k1);
        const double u = u01(r.x, r.y);
        double p = exp(-lam), s = p;
        int k = 0;
        while (u > s && k < 256) { ++k; p *= lam / (double)k; s += p; }
        return (float)k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {
        c.w = attempt;
        const Philox4 r = philox4x32_10(c, key.k0, key.k1);
        const double U = u01(r.x, r.y) - 0.5;
        const double V = u01(r.z, r.w);
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (float)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0)) return (float)kf;
    }
    return (float)floor(lam + 0.5);     // unreachable in practice (acceptance ~ 0.9 per attempt)
}

}  // namespace mvsim
