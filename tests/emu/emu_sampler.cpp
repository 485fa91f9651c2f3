// CPU evaluation of the counter-based Poisson sampler (multiview-simulation_b200/csrc/sampler.cuh is
// __host__ __device__): lets the CPU test suite check Philox known answers and the distribution.
#include "sampler.cuh"

using namespace mvsim;

extern "C" void emu_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    Philox4 c = { ctr[0], ctr[1], ctr[2], ctr[3] };
    const Philox4 r = philox4x32_10(c, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

extern "C" void emu_poisson(const double* lam, float* out, uint64_t n, uint64_t seed, uint64_t stream)
{
    const PoissonKey k = make_poisson_key(seed, stream);
    for (uint64_t g = 0; 4 * g < n; ++g) {
        double l[4];
        float o[4];
        for (int i = 0; i < 4; ++i) l[i] = 4 * g + i < n ? lam[4 * g + i] : 0.0;
        poisson_group4(l, g, k, o);
        for (int i = 0; i < 4 && 4 * g + i < n; ++i) out[4 * g + i] = o[i];
    }
}

// one variate from given random words (top-of-range words probe the float32 CDF plateau of the inversion search);
// returns 1 when the fast path finished it
extern "C" int emu_poisson_fast(double lam, uint32_t ru, uint32_t rv, float* out) { return poisson_fast(lam, ru, rv, *out) ? 1 : 0; }
extern "C" float emu_u01f(uint32_t x) { return u01f(x); }
