// Launch interface of the line-FFT kernels.  The kernels are instantiated per supported size in
// fft_group.cu, compiled once per size group (-DMVSIM_GROUP=k) so the build parallelises.
#pragma once
#include <cuda_runtime.h>
#include "conv_plan.h"
#include "line_fft.cuh"

namespace mvsim {

enum FftKind { FFT_XFWD = 0, FFT_XINV = 1, FFT_SFWD = 2, FFT_SINV = 3, FFT_ZFUSED = 4 };

constexpr int kStridedLanes = 8;     // T: neighbouring kx columns per CTA (64-byte segments)
constexpr int kXThreadsTarget = 256; // x passes: rows per CTA = kXThreadsTarget / threads-per-line

constexpr int x_rows_per_block(int a, int b) { return (kXThreadsTarget / (a > b ? a : b)) > 0 ? kXThreadsTarget / (a > b ? a : b) : 1; }

// returns cudaError_t as int, or -1 when `n` is not in this group
int fft_launch_g0(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s);
int fft_launch_g1(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s);
int fft_launch_g2(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s);
int fft_launch_g3(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s);
int fft_launch_g4(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s);

inline int fft_launch(int kind, int n, const void* params, unsigned gx, unsigned gy, cudaStream_t s)
{
    int r = fft_launch_g0(kind, n, params, gx, gy, s);
    if (r == -1) r = fft_launch_g1(kind, n, params, gx, gy, s);
    if (r == -1) r = fft_launch_g2(kind, n, params, gx, gy, s);
    if (r == -1) r = fft_launch_g3(kind, n, params, gx, gy, s);
    if (r == -1) r = fft_launch_g4(kind, n, params, gx, gy, s);
    return r;
}

}  // namespace mvsim
