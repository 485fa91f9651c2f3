// Host-side probe for the uint16 count transport: how fast can the host cores widen uint16 counts to the
// float32 the reference API returns?  gcc -O3 -mavx2 -fopenmp tools/widen_probe.c -o /tmp/widen_probe && /tmp/widen_probe
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void widen(const uint16_t* src, float* dst, size_t n)
{
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(src + i));
        const __m256 lo = _mm256_cvtepi32_ps(_mm256_cvtepu16_epi32(_mm256_castsi256_si128(v)));
        const __m256 hi = _mm256_cvtepi32_ps(_mm256_cvtepu16_epi32(_mm256_extracti128_si256(v, 1)));
        _mm256_stream_ps(dst + i, lo);
        _mm256_stream_ps(dst + i + 8, hi);
    }
    for (; i < n; ++i) dst[i] = (float)src[i];
}

int main(void)
{
    const size_t n = (size_t)103 * 1024 * 1024;     // one config-3 view
    uint16_t* src = aligned_alloc(64, n * 2);
    float* dst = aligned_alloc(64, n * 4);
    for (size_t i = 0; i < n; ++i) src[i] = (uint16_t)(i * 2654435761u >> 20);
    memset(dst, 0, n * 4);
    const int maxt = omp_get_num_procs();
    for (int t = 1; t <= maxt; t *= 2) {
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            const double t0 = omp_get_wtime();
#pragma omp parallel num_threads(t)
            {
                const int k = omp_get_thread_num(), nt = omp_get_num_threads();
                const size_t chunk = (n / nt + 15) / 16 * 16, a = (size_t)k * chunk, b = a + chunk < n ? a + chunk : n;
                if (a < n) widen(src + a, dst + a, b - a);
            }
            const double dt = omp_get_wtime() - t0;
            if (dt < best) best = dt;
        }
        printf("threads %2d: %.1f ms per view (%.1f GB/s of float32 written)\n", t, best * 1e3, n * 4 / best / 1e9);
    }
    double chk = 0;
    for (size_t i = 0; i < n; i += 4097) chk += dst[i];
    printf("checksum %.0f, processors %d\n", chk, maxt);
    return 0;
}
