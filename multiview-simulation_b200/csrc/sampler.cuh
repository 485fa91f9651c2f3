// Counter-based Poisson sampler: Philox4x32-10 + exact Poisson variates.
//
// Replaces Tools.poissonProcess / PoissonGenerator.nextValue (S/Tools.java:73-86,
// S/uncommons/PoissonGenerator.java:95-109): the reference counts exponential inter-arrival times
// drawn from ONE sequential java.util.Random (O(lambda) Math.log calls per voxel, inherently
// serial).  Here every output voxel owns a Philox counter (key = seed, counter = voxel index,
// stream id, attempt), so the result does not depend on the launch geometry or on how views are
// sharded over GPUs.  The variate is exact Poisson(lambda):
//   lambda < 10 : inversion by sequential search on one uniform,
//   lambda >= 10: PTRS transformed rejection (W. Hoermann, Insurance: Mathematics and Economics 12, 1993),
//   lambda > 1e7: normal limit (counts no longer representable exactly in the float32 output anyway).
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

// (0,1) float uniform from 32 random bits (24 significant)
MVSIM_HD float u01f(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 0x1.0p-24f; }

// One variate from two 32-bit words (ru: inversion uniform / PTRS U, rv: PTRS V).  The arithmetic is
// float32 on the paths nearly every voxel takes (probabilities accurate to ~1e-7, far below what any
// finite sample can resolve); only the rare exact acceptance test of PTRS, whose two sides nearly
// cancel, is evaluated in double.  Rejected PTRS proposals draw fresh words from the voxel's own
// counter (index, attempt >= 2), so the result depends on (seed, stream, voxel index) only.
MVSIM_HD float poisson_one(double lam_d, uint32_t ru, uint32_t rv, uint64_t index, PoissonKey key)
{
    if (!(lam_d > 0.0)) return 0.f;
    if (lam_d < 10.0) {
        // inversion by sequential search: k = min { k : u <= sum_{j<=k} e^-lam lam^j / j! }
        const float lam = (float)lam_d;
        const float u = u01f(ru);
        float p = expf(-lam), s = p;
        int k = 0;
        while (u > s && k < 64) { ++k; p *= lam / (float)k; s += p; }
        return (float)k;
    }
    if (lam_d > 1.0e7) {
        // beyond the float32 resolution of the PTRS proposal (and of the float32 output): normal limit
        const double u1 = ((double)ru + 0.5) * 0x1.0p-32, u2 = ((double)rv + 0.5) * 0x1.0p-32;
        const double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        return (float)floor(lam_d + sqrt(lam_d) * g + 0.5);
    }
    const float lam = (float)lam_d;
    const float slam = sqrtf(lam);
    const float b = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * b;
    const float vr = 0.9277f - 3.6224f / (b - 2.0f);
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {
        if (attempt > 0) {
            Philox4 c;
            c.x = (uint32_t)index; c.y = (uint32_t)(index >> 32); c.z = key.stream_lo; c.w = attempt + 1;
            const Philox4 r = philox4x32_10(c, key.k0, key.k1);
            ru = r.x; rv = r.y;
        }
        const float U = u01f(ru) - 0.5f;
        const float V = u01f(rv);
        const float us = 0.5f - fabsf(U);
        const float kf = floorf((2.0f * a / us + b) * U + lam + 0.43f);
        if (us >= 0.07f && V <= vr) return kf;
        if (kf < 0.0f || (us < 0.013f && V > us)) continue;
        // exact acceptance test  log(V * invalpha / (a/us^2 + b)) <= -lam + k log(lam) - log(k!)
        // with log(k!) by Stirling's series for k >= 10; rewritten around d = k - lam so that the large terms
        // cancel analytically:  rhs = d - k log1p(d/lam) - log(2 pi k)/2 - (1/(12k) - 1/(360k^3) + ...)
        const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
        if (lam <= 3.0e4f) {
            const float lhs = logf(V * invalpha / (a / (us * us) + b));
            float rhs;
            if (kf < 10.0f) {
                rhs = -lam + kf * logf(lam) - lgammaf(kf + 1.0f);
            } else {
                const float d = kf - lam, ik = 1.0f / kf;
                rhs = d - kf * log1pf(d / lam) - 0.5f * logf(6.2831853f * kf) - ik * (0.083333333f - 0.0027777778f * ik * ik);
            }
            if (lhs <= rhs) return kf;
        } else {
            const double k = (double)kf, usd = (double)us, d = k - lam_d, ik = 1.0 / k;
            const double lhs = log((double)V * (double)invalpha / ((double)a / (usd * usd) + (double)b));
            const double rhs = d - k * log1p(d / lam_d) - 0.5 * log(6.283185307179586 * k) -
                               ik * (1.0 / 12.0 - ik * ik * (1.0 / 360.0 - ik * ik / 1260.0));
            if (lhs <= rhs) return kf;
        }
    }
    return floorf(lam + 0.5f);     // unreachable in practice (acceptance ~ 0.9 per attempt)
}

// Four consecutive voxels (flat indices 4g .. 4g+3) share two Philox blocks: block (g, 0) supplies the
// first word of each voxel, block (g, 1) -- generated only if some lambda needs it -- the second.
MVSIM_HD void poisson_group4(const double (&lam)[4], uint64_t group, PoissonKey key, float (&out)[4])
{
    Philox4 c;
    c.x = (uint32_t)group; c.y = (uint32_t)(group >> 32); c.z = key.stream_lo; c.w = 0;
    const Philox4 r0 = philox4x32_10(c, key.k0, key.k1);
    Philox4 r1 = { 0u, 0u, 0u, 0u };
    if (lam[0] >= 10.0 || lam[1] >= 10.0 || lam[2] >= 10.0 || lam[3] >= 10.0) {
        c.w = 1;
        r1 = philox4x32_10(c, key.k0, key.k1);
    }
    out[0] = poisson_one(lam[0], r0.x, r1.x, 4 * group + 0, key);
    out[1] = poisson_one(lam[1], r0.y, r1.y, 4 * group + 1, key);
    out[2] = poisson_one(lam[2], r0.z, r1.z, 4 * group + 2, key);
    out[3] = poisson_one(lam[3], r0.w, r1.w, 4 * group + 3, key);
}

}  // namespace mvsim
