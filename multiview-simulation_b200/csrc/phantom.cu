// Input generators either side of the acquisition path, S = src/main/java/net/preibisch/simulation:
//   bead phantom     S/SimulateBeads.java:97-205   renderPoints / addGaussian (analytic Gaussians x 1000, summed in point order)
//   sphere phantom   S/SimulateMultiViewDataset.java:366-522   drawSpheres (max-composited small spheres) + downSample2x
//   makeSquare       S/Tools.java:315-349          centre-pad to a cube with the minimum
// The reference walks ImgLib2 cursors on one thread.  Here the only sequential part that remains is the replay of
// java.util.Random (a 48-bit LCG whose draws are position dependent), done on the host by the callers in capi.cu;
// all voxel work is CUDA.
#include <math.h>

#include <algorithm>
#include <map>

#include "ctx.h"

namespace mvsim {

static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

#define MVSIM_LAUNCH_CHECK(ctx)                                                     \
    do {                                                                            \
        (ctx)->launches++;                                                          \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, "kernel launch");      \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Beads.  addGaussian (:168-205) adds (float)(ex*ey*ez) * 1000f to every voxel of a box of 2*diameter voxels per axis
// around round(p); the float sums depend on the ORDER of the points, so a scatter with atomics would not reproduce the
// reference.  Gather instead: the host bins the points into 32x8x8 bricks (lists stay in point order), one CTA per
// non-empty brick walks its list; per (brick, bead) the 48 one-dimensional exponentials are evaluated once in FP64 and
// shared, each thread then accumulates its 8 voxels in FP32 in the reference's operation order.
// ---------------------------------------------------------------------------------------------
constexpr int BX = 32, BY = 8, BZ = 8;

struct BeadRec { double p[3]; int mn[3]; int pad; };

__global__ void __launch_bounds__(256) beads_brick_kernel(float* __restrict__ out, int X, int Y, int Z, const int4* __restrict__ bricks,
                                                          const int* __restrict__ list, const BeadRec* __restrict__ beads,
                                                          int sx, int sy, int sz, double tx, double ty, double tz)
{
    __shared__ double e[BX + BY + BZ];
    __shared__ unsigned char in_box[BX + BY + BZ];
    const int4 b = bricks[blockIdx.x];              // brick coordinates, w = first list entry; the next brick's w ends the list
    const int end = bricks[blockIdx.x + 1].w;
    const int x0 = b.x * BX, y0 = b.y * BY, z0 = b.z * BZ;
    const int tx_ = threadIdx.x & (BX - 1), ty_ = threadIdx.x / BX;
    float acc[BZ];
#pragma unroll
    for (int k = 0; k < BZ; ++k) acc[k] = 0.f;
    for (int it = b.w; it < end; ++it) {
        const BeadRec r = beads[list[it]];
        __syncthreads();
        if (threadIdx.x < BX + BY + BZ) {
            const int t = threadIdx.x;
            const int d = t < BX ? 0 : (t < BX + BY ? 1 : 2);
            const int c = d == 0 ? x0 + t : (d == 1 ? y0 + t - BX : z0 + t - BX - BY);
            const int size = d == 0 ? sx : (d == 1 ? sy : sz);
            const double two_sq = d == 0 ? tx : (d == 1 ? ty : tz);
            const double dx = __dsub_rn(r.p[d], (double)c);
            e[t] = exp(__ddiv_rn(-__dmul_rn(dx, dx), two_sq));
            in_box[t] = (c >= r.mn[d] && c < r.mn[d] + size) ? 1 : 0;
        }
        __syncthreads();
        if (in_box[tx_] && in_box[BX + ty_]) {
            const double exy = __dmul_rn(e[tx_], e[BX + ty_]);              // value = 1; value *= ex; value *= ey; value *= ez
#pragma unroll
            for (int k = 0; k < BZ; ++k)
                if (in_box[BX + BY + k]) acc[k] = __fadd_rn(acc[k], __fmul_rn((float)__dmul_rn(exy, e[BX + BY + k]), 1000.0f));
        }
    }
    const int x = x0 + tx_, y = y0 + ty_;
    if (x < X && y < Y)
#pragma unroll
        for (int k = 0; k < BZ; ++k)
            if (z0 + k < Z) out[x + (long long)X * (y + (long long)Y * (z0 + k))] = acc[k];
}

// imglib2 Util.getSuggestedKernelDiameter
static int suggested_kernel_diameter(double sigma)
{
    int size = 3;
    if (sigma > 0) size = std::max(3, 2 * (int)(3.0 * sigma + 0.5) + 1);
    return size;
}

int k_render_beads(mvsim_ctx* ctx, const double* points, int n, const double sigma[3], const int64_t imin[3], const int64_t imax[3], float* d_out)
{
    const int64_t X = imax[0] - imin[0], Y = imax[1] - imin[1], Z = imax[2] - imin[2];       // :106 (max - min, not dimension)
    const size_t bytes = (size_t)(X * Y * Z) * sizeof(float);
    MVSIM_CUDA(ctx, cudaMemsetAsync(d_out, 0, bytes, ctx->stream));
    int size[3];
    double two_sq[3];
    for (int d = 0; d < 3; ++d) { size[d] = suggested_kernel_diameter(sigma[d]) * 2; two_sq[d] = 2 * sigma[d] * sigma[d]; }
    std::vector<BeadRec> beads;
    std::map<int64_t, std::vector<int>> bins;           // brick id (z major) -> bead indices in point order
    const int64_t nbx = (X + BX - 1) / BX, nby = (Y + BY - 1) / BY, nbz = (Z + BZ - 1) / BZ;
    for (int i = 0; i < n; ++i) {
        BeadRec r;
        bool inside = true;
        for (int d = 0; d < 3 && inside; ++d) {
            r.p[d] = points[3 * i + d] - (double)imin[d];                                    // isInsideAdjust (:123-135)
            if (!(r.p[d] >= 0 && r.p[d] <= (double)(imax[d] - imin[d]))) inside = false;     // <= dimension - 1
        }
        if (!inside) continue;
        int64_t lo[3], hi[3];
        const int64_t dim[3] = { X, Y, Z };
        bool visible = true;
        for (int d = 0; d < 3; ++d) {
            r.mn[d] = (int)floor(r.p[d] + 0.5) - size[d] / 2;                                // (int)Math.round(location) - size/2
            lo[d] = std::max<int64_t>(r.mn[d], 0);
            hi[d] = std::min<int64_t>((int64_t)r.mn[d] + size[d] - 1, dim[d] - 1);
            if (lo[d] > hi[d]) visible = false;
        }
        r.pad = 0;
        if (!visible) continue;
        const int id = (int)beads.size();
        beads.push_back(r);
        for (int64_t bz = lo[2] / BZ; bz <= hi[2] / BZ; ++bz)
            for (int64_t by = lo[1] / BY; by <= hi[1] / BY; ++by)
                for (int64_t bx = lo[0] / BX; bx <= hi[0] / BX; ++bx) bins[bx + nbx * (by + nby * bz)].push_back(id);
    }
    (void)nbz;
    if (bins.empty()) return MVSIM_OK;
    std::vector<int4> bricks;
    std::vector<int> list;
    for (auto& kv : bins) {
        const int64_t id = kv.first;
        bricks.push_back(make_int4((int)(id % nbx), (int)((id / nbx) % nby), (int)(id / (nbx * nby)), (int)list.size()));
        list.insert(list.end(), kv.second.begin(), kv.second.end());
    }
    bricks.push_back(make_int4(0, 0, 0, (int)list.size()));
    void *d_bricks = nullptr, *d_list = nullptr, *d_beads = nullptr;
    int st = dev_alloc(ctx, &d_bricks, bricks.size() * sizeof(int4));
    if (st == MVSIM_OK) st = dev_alloc(ctx, &d_list, list.size() * sizeof(int));
    if (st == MVSIM_OK) st = dev_alloc(ctx, &d_beads, beads.size() * sizeof(BeadRec));
    cudaError_t e = cudaSuccess;
    if (st == MVSIM_OK) {
        // the host vectors die at return: synchronous copies (tables are a few hundred KB)
        e = cudaMemcpyAsync(d_bricks, bricks.data(), bricks.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_list, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_beads, beads.data(), beads.size() * sizeof(BeadRec), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            beads_brick_kernel<<<(unsigned)(bricks.size() - 1), 256, 0, ctx->stream>>>(d_out, (int)X, (int)Y, (int)Z, (const int4*)d_bricks, (const int*)d_list,
                                                                                     (const BeadRec*)d_beads, size[0], size[1], size[2], two_sq[0], two_sq[1], two_sq[2]);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    dev_free(ctx, d_bricks);
    dev_free(ctx, d_list);
    dev_free(ctx, d_beads);
    if (st != MVSIM_OK) return st;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "render_beads");
    return MVSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// Sphere phantom.  drawSpheres paints each small sphere with max(value, existing) (:517): the result does not depend
// on the painting order, so the spheres are painted concurrently with an integer atomicMax on the float bits (all values
// are >= 0, where float order == int order).  Membership follows imglib2's HyperSphereCursor: NESTED integer radii,
// ry = (long)sqrt(r^2 - dz^2), rx = (long)sqrt(ry^2 - dy^2).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int isqrt_floor(int v)           // v <= 2^22 here; (long)Math.sqrt of an exact integer
{
    int r = (int)sqrtf((float)v);
    while (r * r > v) --r;
    while ((r + 1) * (r + 1) <= v) ++r;
    return r;
}

// one CTA per small sphere; rec = (cx, cy, cz, radius), value
__global__ void __launch_bounds__(256) paint_spheres_kernel(float* __restrict__ img, int X, int Y, int Z, const int4* __restrict__ rec,
                                                            const float* __restrict__ val)
{
    const int4 s = rec[blockIdx.x];
    const int R = s.w, D = 2 * R + 1;
    const int bits = __float_as_int(val[blockIdx.x]);
    for (int row = threadIdx.x / 32; row < D * D; row += blockDim.x / 32) {       // a warp per (dz, dy) row
        const int dz = row / D - R, dy = row % D - R;
        const int ry = isqrt_floor(R * R - dz * dz);
        if (dy < -ry || dy > ry) continue;
        const int rx = isqrt_floor(ry * ry - dy * dy);
        const int z = s.z + dz, y = s.y + dy;
        if ((unsigned)z >= (unsigned)Z || (unsigned)y >= (unsigned)Y) continue;
        int* line = reinterpret_cast<int*>(img) + (long long)X * (y + (long long)Y * z);
        for (int dx = -rx + (int)(threadIdx.x & 31); dx <= rx; dx += 32) {
            const int x = s.x + dx;
            if ((unsigned)x < (unsigned)X) atomicMax(line + x, bits);
        }
    }
}

int k_paint_spheres(mvsim_ctx* ctx, float* d_img, const int64_t dims[3], const int* host_rec /* n * 4 */, const float* host_val, int n)
{
    const size_t bytes = (size_t)(dims[0] * dims[1] * dims[2]) * sizeof(float);
    MVSIM_CUDA(ctx, cudaMemsetAsync(d_img, 0, bytes, ctx->stream));
    if (n == 0) return MVSIM_OK;
    void *d_rec = nullptr, *d_val = nullptr;
    int st = dev_alloc(ctx, &d_rec, (size_t)n * sizeof(int4));
    if (st == MVSIM_OK) st = dev_alloc(ctx, &d_val, (size_t)n * sizeof(float));
    cudaError_t e = cudaSuccess;
    if (st == MVSIM_OK) {
        e = cudaMemcpyAsync(d_rec, host_rec, (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_val, host_val, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            paint_spheres_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(d_img, (int)dims[0], (int)dims[1], (int)dims[2], (const int4*)d_rec, (const float*)d_val);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);       // host_rec / host_val belong to the caller
    }
    dev_free(ctx, d_rec);
    dev_free(ctx, d_val);
    if (st != MVSIM_OK) return st;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "paint_spheres");
    return MVSIM_OK;
}

// downSample2x (:394-423): out dims = dims/2 - 1, sample at 2l + 0.5.  floor = 2l, all fractional weights are exactly
// 0.5 and 2l + 1 <= dims - 3, so no mirror tap is ever taken; imglib2's blend is f32(tap * 0.125) summed in the tap
// order 000,100,110,010,011,111,101,001 (x,y,z bits).  Two output voxels per thread, float4 row loads.
__global__ void __launch_bounds__(256) downsample2x_kernel(const float* __restrict__ in, float* __restrict__ out, int X, int Y, int OX, int OY, int OZ)
{
    const int OXP = (OX + 1) / 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)OXP * OY * OZ) return;
    const int xp = (int)(i % OXP);
    const long long r = i / OXP;
    const int y = (int)(r % OY), z = (int)(r / OY);
    const long long base = 4ll * xp + (long long)X * (2 * y + (long long)Y * (2 * z));
    const long long sy = X, sz = (long long)X * Y;
    float t[4][4];      // [row: 00, 10 (y+1), 11 (y+1,z+1), 01 (z+1)][x .. x+3]
    const long long off[4] = { 0, sy, sy + sz, sz };
    const bool vec = (X % 4 == 0) && (4 * xp + 3 < X);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (vec) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(in + base + off[k]));
            t[k][0] = v.x; t[k][1] = v.y; t[k][2] = v.z; t[k][3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) t[k][j] = (4 * xp + j < X) ? __ldg(in + base + off[k] + j) : 0.f;
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int x = 2 * xp + h;
        if (x >= OX) break;
        const int a = 2 * h, b = 2 * h + 1;
        float acc = __fmul_rn(t[0][a], 0.125f);                 // 000
        acc = __fadd_rn(acc, __fmul_rn(t[0][b], 0.125f));       // 100
        acc = __fadd_rn(acc, __fmul_rn(t[1][b], 0.125f));       // 110
        acc = __fadd_rn(acc, __fmul_rn(t[1][a], 0.125f));       // 010
        acc = __fadd_rn(acc, __fmul_rn(t[2][a], 0.125f));       // 011
        acc = __fadd_rn(acc, __fmul_rn(t[2][b], 0.125f));       // 111
        acc = __fadd_rn(acc, __fmul_rn(t[3][b], 0.125f));       // 101
        acc = __fadd_rn(acc, __fmul_rn(t[3][a], 0.125f));       // 001
        out[x + (long long)OX * (y + (long long)OY * z)] = acc;
    }
}

int k_downsample2x(mvsim_ctx* ctx, const float* in, const int64_t dims[3], float* out)
{
    const int OX = (int)(dims[0] / 2 - 1), OY = (int)(dims[1] / 2 - 1), OZ = (int)(dims[2] / 2 - 1);
    const size_t n = (size_t)((OX + 1) / 2) * OY * OZ;
    downsample2x_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(in, out, (int)dims[0], (int)dims[1], OX, OY, OZ);
    MVSIM_LAUNCH_CHECK(ctx);
    return MVSIM_OK;
}

// ---------------------------------------------------------------------------------------------
// makeSquare (S/Tools.java:315-349): minimum of the input, then a cube of the largest dimension with the input centred
// at offset square/2 - dim/2 (integer divisions) and the minimum elsewhere.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ordered_bits(float f)       // monotone float -> unsigned map
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(256) min_kernel(const float* __restrict__ in, size_t n, unsigned* __restrict__ result)
{
    unsigned m = 0xffffffffu;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = in[i];
        if (v == v) m = min(m, ordered_bits(v));        // Math.min would propagate NaN; a PSF holds none
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(result, m);
}

__global__ void __launch_bounds__(256) make_square_kernel(const float* __restrict__ in, float* __restrict__ out, int X, int Y, int Z, int M,
                                                          const unsigned* __restrict__ min_bits)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * M * M) return;
    const int x = (int)(i % M), y = (int)((i / M) % M), z = (int)(i / ((long long)M * M));
    const int sx = x - M / 2 + X / 2, sy = y - M / 2 + Y / 2, sz = z - M / 2 + Z / 2;
    const bool inside = (unsigned)sx < (unsigned)X && (unsigned)sy < (unsigned)Y && (unsigned)sz < (unsigned)Z;
    out[i] = inside ? __ldg(in + sx + (long long)X * (sy + (long long)Y * sz)) : from_ordered_bits(*min_bits);
}

int k_make_square(mvsim_ctx* ctx, const float* in, const int64_t dims[3], float* out)
{
    const int M = (int)std::max(dims[0], std::max(dims[1], dims[2]));
    const size_t n = (size_t)(dims[0] * dims[1] * dims[2]);
    unsigned* d_min = reinterpret_cast<unsigned*>(ctx->d_scalars + 7);
    MVSIM_CUDA(ctx, cudaMemsetAsync(d_min, 0xff, sizeof(unsigned), ctx->stream));
    min_kernel<<<std::min<unsigned>(blocks_for(n, 256), 148 * 8), 256, 0, ctx->stream>>>(in, n, d_min);
    MVSIM_LAUNCH_CHECK(ctx);
    make_square_kernel<<<blocks_for((size_t)M * M * M, 256), 256, 0, ctx->stream>>>(in, out, (int)dims[0], (int)dims[1], (int)dims[2], M, d_min);
    MVSIM_LAUNCH_CHECK(ctx);
    return MVSIM_OK;
}

}  // namespace mvsim
