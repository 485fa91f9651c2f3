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

// Approximate float division / square root / exp / log on the paths where a relative error of ~1e-7 only perturbs a
// probability by as much (MUFU on the device, libm on the host emulation)
MVSIM_HD float fast_div(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
MVSIM_HD float fast_sqrt(float a)
{
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return sqrtf(a);
#endif
}
MVSIM_HD float fast_exp(float a)
{
#ifdef __CUDA_ARCH__
    return __expf(a);
#else
    return expf(a);
#endif
}
MVSIM_HD float fast_log(float a)
{
#ifdef __CUDA_ARCH__
    return __logf(a);
#else
    return logf(a);
#endif
}

// float uniform strictly inside (0,1) from the top 23 of 32 random bits: (m + 1/2) 2^-23, m < 2^23.  Every value
// (2m + 1) 2^-24 <= 1 - 2^-24 is exactly representable; with 24 bits the "+ 0.5f" of m = 2^24 - 1 rounded to 2^24, i.e. u = 1.0f
// with probability 2^-24 per draw.
MVSIM_HD float u01f(uint32_t x) { return ((float)(x >> 9) + 0.5f) * 0x1.0p-23f; }

// PTRS hat-function constants for one lambda (float32 everywhere on the fast path)
struct PtrsParams { float lam, b, a, vr; };

MVSIM_HD PtrsParams ptrs_params(float lam)
{
    PtrsParams q;
    q.lam = lam;
    q.b = 0.931f + 2.53f * fast_sqrt(lam);
    q.a = -0.059f + 0.02483f * q.b;
    q.vr = 0.9277f - fast_div(3.6224f, q.b - 2.0f);
    return q;
}

// One PTRS proposal from two random words.  Returns 1 = accepted by the squeeze (k valid), 0 = rejected outright,
// 2 = needs the exact acceptance test (k, us, V valid).
MVSIM_HD int ptrs_propose(const PtrsParams& q, uint32_t ru, uint32_t rv, float& kf, float& us, float& V)
{
    const float U = u01f(ru) - 0.5f;
    V = u01f(rv);
    us = 0.5f - fabsf(U);
    kf = floorf((fast_div(2.0f * q.a, us) + q.b) * U + q.lam + 0.43f);
    if (us >= 0.07f && V <= q.vr) return 1;
    if (kf < 0.0f || (us < 0.013f && V > us)) return 0;
    return 2;
}

// exact acceptance test  log(V * invalpha / (a/us^2 + b)) <= -lam + k log(lam) - log(k!)
// with log(k!) by Stirling's series for k >= 10; rewritten around d = k - lam so that the large terms
// cancel analytically:  rhs = d - k log1p(d/lam) - log(2 pi k)/2 - (1/(12k) - 1/(360k^3) + ...)
MVSIM_HD bool ptrs_accept(const PtrsParams& q, double lam_d, float kf, float us, float V)
{
    const float invalpha = 1.1239f + fast_div(1.1328f, q.b - 3.4f);
    if (q.lam <= 3.0e4f) {
        const float lhs = fast_log(fast_div(V * invalpha, fast_div(q.a, us * us) + q.b));
        float rhs;
        if (kf < 10.0f) {
            rhs = -q.lam + kf * logf(q.lam) - lgammaf(kf + 1.0f);
        } else {
            const float d = kf - q.lam, ik = 1.0f / kf;
            rhs = d - kf * log1pf(d / q.lam) - 0.5f * logf(6.2831853f * kf) - ik * (0.083333333f - 0.0027777778f * ik * ik);
        }
        return lhs <= rhs;
    }
    const double k = (double)kf, usd = (double)us, d = k - lam_d, ik = 1.0 / k;
    const double lhs = log((double)V * (double)invalpha / ((double)q.a / (usd * usd) + (double)q.b));
    const double rhs = d - k * log1p(d / lam_d) - 0.5 * log(6.283185307179586 * k) - ik * (1.0 / 12.0 - ik * ik * (1.0 / 360.0 - ik * ik / 1260.0));
    return lhs <= rhs;
}

// One attempt at a voxel whose first proposal (ru, rv) was not accepted by the squeeze.  Attempt 0 applies the exact test
// to that first proposal; attempt >= 1 draws a fresh proposal from the voxel's own counter (index, attempt + 1).
// Returns true when a variate was accepted (kf valid).
MVSIM_HD bool ptrs_attempt(const PtrsParams& q, double lam_d, uint32_t ru, uint32_t rv, uint64_t index, PoissonKey key, uint32_t attempt, float& kf)
{
    if (attempt > 0) {
        Philox4 c;
        c.x = (uint32_t)index; c.y = (uint32_t)(index >> 32); c.z = key.stream_lo; c.w = attempt + 1;
        const Philox4 r = philox4x32_10(c, key.k0, key.k1);
        ru = r.x; rv = r.y;
    }
    float us, V;
    const int st = ptrs_propose(q, ru, rv, kf, us, V);
    return st == 1 || (st == 2 && ptrs_accept(q, lam_d, kf, us, V));
}

constexpr uint32_t kPtrsMaxAttempts = 64;      // unreachable in practice (acceptance ~ 0.9 per proposal)
MVSIM_HD float ptrs_give_up(double lam_d) { return floorf((float)lam_d + 0.5f); }

// Finishes a voxel by itself: attempts until one is accepted.
MVSIM_HD float ptrs_resolve(double lam_d, uint32_t ru, uint32_t rv, uint64_t index, PoissonKey key)
{
    const PtrsParams q = ptrs_params((float)lam_d);
    float kf;
    for (uint32_t attempt = 0; attempt < kPtrsMaxAttempts; ++attempt)
        if (ptrs_attempt(q, lam_d, ru, rv, index, key, attempt, kf)) return kf;
    return ptrs_give_up(lam_d);
}

// Fast part of one variate.  Returns true when done (out valid); false when the voxel needs ptrs_resolve.
// The arithmetic is float32 on the paths nearly every voxel takes (probabilities accurate to ~1e-7, far below what
// any finite sample can resolve).  The result depends on (seed, stream, voxel index) only.
MVSIM_HD bool poisson_fast(double lam_d, uint32_t ru, uint32_t rv, float& out)
{
    out = 0.f;
    if (!(lam_d > 0.0)) return true;
    if (lam_d < 10.0) {
        // inversion by sequential search: k = min { k : u <= sum_{j<=k} e^-lam lam^j / j! }
        const float lam = (float)lam_d;
        const float u = u01f(ru);
        // The float32 CDF saturates just below 1 (0.9999998 .. 0.99999994 depending on lam) once p < ulp(s)/2: a u above that
        // plateau ends the search at the first k whose term no longer changes the sum (the exact quantile lies within a
        // couple of counts of it; such u have probability < 2e-7) instead of running on to the iteration cap.
        float p = fast_exp(-lam), s = p;
        int k = 0;
        while (u > s && k < 64) {
            ++k;
            p *= fast_div(lam, (float)k);
            const float s2 = s + p;
            if (s2 == s) break;
            s = s2;
        }
        out = (float)k;
        return true;
    }
    if (lam_d > 1.0e7) {
        // beyond the float32 resolution of the PTRS proposal (and of the float32 output): normal limit
        const double u1 = ((double)ru + 0.5) * 0x1.0p-32, u2 = ((double)rv + 0.5) * 0x1.0p-32;
        const double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        out = (float)floor(lam_d + sqrt(lam_d) * g + 0.5);
        return true;
    }
    float us, V;
    return ptrs_propose(ptrs_params((float)lam_d), ru, rv, out, us, V) == 1;
}

// Four consecutive voxels (flat indices 4g .. 4g+3) share two Philox blocks: block (g, 0) supplies the first word
// of each voxel, block (g, 1) -- generated only if some lambda needs it -- the second.  Voxels whose first PTRS
// proposal fails the squeeze are finished afterwards in ONE loop (one copy of the slow code, and lanes of a warp
// share its iterations), instead of diverging four times.
// fast part for the four voxels of a group: returns the 4-bit mask of voxels that still need ptrs_resolve and the random
// words of the group (the first proposal of a pending voxel is re-derived from them)
MVSIM_HD unsigned poisson_group4_fast(const double (&lam)[4], uint64_t group, PoissonKey key, float (&out)[4], Philox4& r0, Philox4& r1)
{
    Philox4 c;
    c.x = (uint32_t)group; c.y = (uint32_t)(group >> 32); c.z = key.stream_lo; c.w = 0;
    r0 = philox4x32_10(c, key.k0, key.k1);
    r1.x = r1.y = r1.z = r1.w = 0u;
    if (lam[0] >= 10.0 || lam[1] >= 10.0 || lam[2] >= 10.0 || lam[3] >= 10.0) {
        c.w = 1;
        r1 = philox4x32_10(c, key.k0, key.k1);
    }
    unsigned pending = 0;
    if (!poisson_fast(lam[0], r0.x, r1.x, out[0])) pending |= 1u;
    if (!poisson_fast(lam[1], r0.y, r1.y, out[1])) pending |= 2u;
    if (!poisson_fast(lam[2], r0.z, r1.z, out[2])) pending |= 4u;
    if (!poisson_fast(lam[3], r0.w, r1.w, out[3])) pending |= 8u;
    return pending;
}

MVSIM_HD void poisson_group4(const double (&lam)[4], uint64_t group, PoissonKey key, float (&out)[4])
{
    Philox4 r0, r1;
    unsigned pending = poisson_group4_fast(lam, group, key, out, r0, r1);
    while (pending) {
        // lowest pending voxel; selects instead of dynamic indexing keep everything in registers
        const int i = (pending & 1u) ? 0 : (pending & 2u) ? 1 : (pending & 4u) ? 2 : 3;
        const double l = i == 0 ? lam[0] : i == 1 ? lam[1] : i == 2 ? lam[2] : lam[3];
        const uint32_t ru = i == 0 ? r0.x : i == 1 ? r0.y : i == 2 ? r0.z : r0.w;
        const uint32_t rv = i == 0 ? r1.x : i == 1 ? r1.y : i == 2 ? r1.z : r1.w;
        const float k = ptrs_resolve(l, ru, rv, 4 * group + (uint64_t)i, key);
        if (i == 0) out[0] = k; else if (i == 1) out[1] = k; else if (i == 2) out[2] = k; else out[3] = k;
        pending &= pending - 1;
    }
}

}  // namespace mvsim
