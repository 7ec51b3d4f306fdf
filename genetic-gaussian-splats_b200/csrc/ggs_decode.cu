// Genome decode for the whole population in one launch.
//
// Restates, per splat and on device, the reference's
//   axes_angle_to_cholesky / genome_to_renderer[_batched]   modules/encode.py:5-24, 28-59, 63-79
//   _preprocess_genome                                     modules/render.py:9-47
// The integer AABB is semantics (splats are hard-clipped to it, render.py:175-177), so this
// file reproduces the reference's fp32 operation order exactly: every arithmetic step is an
// explicitly rounded intrinsic (__fmul_rn/__fadd_rn/... are never contracted into FMAs),
// transcendentals are the IEEE-accurate libdevice expf/logf/sinf/cosf (the functions torch's
// CUDA kernels call), sqrt and division are correctly rounded.  This translation unit is also
// compiled with -fmad=false as a second line of defence.
#include <math.h>

#include "ggs_common.cuh"

namespace ggs {
namespace {

__device__ __forceinline__ float clamp_nan(float v, float lo, float hi)
{
    // torch.clamp: NaN propagates, otherwise min(max(v, lo), hi)
    return (v != v) ? v : fminf(fmaxf(v, lo), hi);
}

__device__ __forceinline__ float clamp_min_nan(float v, float lo)
{
    return (v != v) ? v : fmaxf(v, lo);
}

struct Chol {
    float x, y, log_l11, log_l22, l21, r, g, b, a;
};

// modules/encode.py:5-24 + :35-57.  g = (x, y, log sx, log sy, theta, r, g, b, alpha)
__device__ __forceinline__ Chol encode_axes(const float *g)
{
    const float sx = expf(g[2]);  // encode.py:6
    const float sy = expf(g[3]);  // encode.py:7
    const float c = cosf(g[4]);   // encode.py:8
    const float s = sinf(g[4]);   // encode.py:9
    const float sx2 = __fmul_rn(sx, sx), sy2 = __fmul_rn(sy, sy);
    const float c2 = __fmul_rn(c, c), s2 = __fmul_rn(s, s);
    const float vxx = __fadd_rn(__fmul_rn(sx2, c2), __fmul_rn(sy2, s2));    // encode.py:12
    const float vxy = __fmul_rn(__fmul_rn(__fsub_rn(sx2, sy2), s), c);      // encode.py:13
    const float vyy = __fadd_rn(__fmul_rn(sx2, s2), __fmul_rn(sy2, c2));    // encode.py:14
    const float eps = 1e-12f;                                               // encode.py:16
    const float l11 = __fsqrt_rn(clamp_min_nan(vxx, eps));                  // encode.py:17
    const float l21 = __fdiv_rn(vxy, l11);                                  // encode.py:18
    const float l22 = __fsqrt_rn(clamp_min_nan(__fsub_rn(vyy, __fmul_rn(l21, l21)), eps));  // :19
    Chol o;
    o.x = g[0];
    o.y = g[1];
    o.log_l11 = logf(l11);  // encode.py:21
    o.log_l22 = logf(l22);  // encode.py:22
    o.l21 = l21;            // encode.py:23
    o.r = clamp_nan(g[5], 0.0f, 255.0f);  // encode.py:57,77
    o.g = clamp_nan(g[6], 0.0f, 255.0f);
    o.b = clamp_nan(g[7], 0.0f, 255.0f);
    o.a = clamp_nan(g[8], 0.0f, 255.0f);
    return o;
}

__device__ __forceinline__ Chol load_chol(const float *g)
{
    Chol o;
    o.x = g[0];
    o.y = g[1];
    o.log_l11 = g[2];
    o.log_l22 = g[3];
    o.l21 = g[4];
    o.r = g[5];
    o.g = g[6];
    o.b = g[7];
    o.a = g[8];
    return o;
}

struct Decoded {
    float cx, cy, sxx, sxy, syy, rc, gc, bc, a;
    float hx, hy;  // k-sigma half extents (render.py:24-25)
    int x0, x1, y0, y1;
};

// modules/render.py:14-43
__device__ __forceinline__ Decoded decode_chol(const Chol &g, int H, int W, float k_sigma)
{
    Decoded d;
    const float maxx = (float)(W - 1), maxy = (float)(H - 1);                 // render.py:14
    const float cx = __fmul_rn(clamp_nan(g.x, 0.0f, 1.0f), maxx);             // render.py:15
    const float cy = __fmul_rn(clamp_nan(g.y, 0.0f, 1.0f), maxy);             // render.py:16
    const float l11 = clamp_min_nan(expf(g.log_l11), 1e-6f);                  // render.py:19
    const float l22 = clamp_min_nan(expf(g.log_l22), 1e-6f);                  // render.py:20
    const float l21 = g.l21;                                                  // render.py:21
    const float hx = clamp_min_nan(__fmul_rn(k_sigma, fabsf(l11)), 1.0f);     // render.py:24
    const float hy =
        clamp_min_nan(__fmul_rn(k_sigma, __fadd_rn(fabsf(l21), fabsf(l22))), 1.0f);  // render.py:25
    d.x0 = (int)floorf(clamp_nan(__fsub_rn(cx, hx), 0.0f, maxx));             // render.py:27
    d.x1 = (int)ceilf(clamp_nan(__fadd_rn(cx, hx), 0.0f, maxx));              // render.py:28
    d.y0 = (int)floorf(clamp_nan(__fsub_rn(cy, hy), 0.0f, maxy));             // render.py:29
    d.y1 = (int)ceilf(clamp_nan(__fadd_rn(cy, hy), 0.0f, maxy));              // render.py:30
    const float i11 = __fdiv_rn(1.0f, l11);                                   // render.py:32
    const float i22 = __fdiv_rn(1.0f, l22);                                   // render.py:33
    const float i21 = __fmul_rn(-l21, __fmul_rn(i11, i22));                   // render.py:34
    d.sxx = __fadd_rn(__fmul_rn(i11, i11), __fmul_rn(i21, i21));              // render.py:36
    d.sxy = __fmul_rn(i21, i22);                                              // render.py:37
    d.syy = __fmul_rn(i22, i22);                                              // render.py:38
    d.rc = __fdiv_rn(clamp_nan(g.r, 0.0f, 255.0f), 255.0f);                   // render.py:40
    d.gc = __fdiv_rn(clamp_nan(g.g, 0.0f, 255.0f), 255.0f);                   // render.py:41
    d.bc = __fdiv_rn(clamp_nan(g.b, 0.0f, 255.0f), 255.0f);                   // render.py:42
    d.a = __fdiv_rn(clamp_nan(g.a, 0.0f, 255.0f), 255.0f);                    // render.py:43
    d.cx = cx;
    d.cy = cy;
    d.hx = hx;
    d.hy = hy;
    return d;
}

// Stage this block's rows into shared memory with 128-bit coalesced loads, then hand each
// thread its own row (stride `cols` floats: conflict-free for the usual cols = 9).
__device__ __forceinline__ const float *stage_rows(const float *__restrict__ genomes, int cols,
                                                   int64_t rows, float *sm, float *tmp)
{
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * kDecodeThreads;
    const int nrows = (int)min((int64_t)kDecodeThreads, rows - row0);
    const float *src = genomes + row0 * cols;
    if (cols <= kDecodeStageMaxCols) {
        const int nfl = nrows * cols;
        int done = 0;
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            const int n4 = nfl >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(sm);
            for (int i = tid; i < n4; i += kDecodeThreads) dst4[i] = __ldg(src4 + i);
            done = n4 << 2;
        }
        for (int i = done + tid; i < nfl; i += kDecodeThreads) sm[i] = __ldg(src + i);
        __syncthreads();
        return (tid < nrows) ? sm + tid * cols : nullptr;
    }
    if (tid >= nrows) return nullptr;
#pragma unroll
    for (int k = 0; k < 9; ++k) tmp[k] = __ldg(src + (int64_t)tid * cols + k);
    return tmp;
}

template <bool kAxes>
__global__ void __launch_bounds__(kDecodeThreads)
decode_kernel(const float *__restrict__ genomes, int cols, int64_t rows, int H, int W,
              float k_sigma, float4 *__restrict__ rec, uint2 *__restrict__ aabb,
              float *__restrict__ raw_f, int32_t *__restrict__ raw_i, int *__restrict__ counters,
              int n_counters)
{
    extern __shared__ __align__(16) float sm[];
    pdl_wait();
    pdl_trigger();
    if (blockIdx.x == 0 && counters != nullptr)
        for (int i = threadIdx.x; i < n_counters; i += kDecodeThreads) counters[i] = 0;

    float tmp[9];
    const float *g = stage_rows(genomes, cols, rows, sm, tmp);
    if (g == nullptr) return;
    const int64_t row = (int64_t)blockIdx.x * kDecodeThreads + threadIdx.x;

    const Chol ch = kAxes ? encode_axes(g) : load_chol(g);
    const Decoded d = decode_chol(ch, H, W, k_sigma);

    if (raw_f != nullptr) {
        raw_f[0 * rows + row] = d.cx;
        raw_f[1 * rows + row] = d.cy;
        raw_f[2 * rows + row] = d.sxx;
        raw_f[3 * rows + row] = d.sxy;
        raw_f[4 * rows + row] = d.syy;
        raw_f[5 * rows + row] = d.rc;
        raw_f[6 * rows + row] = d.gc;
        raw_f[7 * rows + row] = d.bc;
        raw_f[8 * rows + row] = d.a;
        raw_i[0 * rows + row] = d.x0;
        raw_i[1 * rows + row] = d.x1;
        raw_i[2 * rows + row] = d.y0;
        raw_i[3 * rows + row] = d.y1;
    }
    if (rec != nullptr) {
        // f = exp(-0.5*quad)*a  ==  2^(A qx^2 + Bq qx qy + Cq qy^2 + log2 a)
        const float kHalfLog2e = 0.72134752044448170368f;
        const float kLog2e = 1.44269504088896340736f;
        SplatRec r;
        r.cx = d.cx;
        r.cy = d.cy;
        r.A = -kHalfLog2e * d.sxx;
        r.Bq = -kLog2e * d.sxy;
        r.Cq = -kHalfLog2e * d.syy;
        r.la = log2f(d.a);  // alpha 0 -> -inf -> f = 0
        r.r = d.rc;
        r.g = d.gc;
        r.b = d.bc;
        r.xpack = (d.x0 & 0xffff) | (d.x1 << 16);
        r.ypack = (d.y0 & 0xffff) | (d.y1 << 16);
        // Column recurrence of the raster (f(i+2) = f(i)*g(i), g(i+2) = g(i)*h): h = 2^(8*Cq).
        // It is used only while the exponent moves by < 64 across a thread's strip of rows
        // anywhere a lane of the tile can sit (|qy| <= hy+1, |qx| <= hx+32); otherwise the
        // splat is marked steep (h = -1) and takes the exact per-pixel path.
        constexpr float kSpan = (float)(kRowsPerThread - 1);  // rows crossed by one strip
        const float swing = fabsf(r.Cq) * (2.0f * kSpan * (d.hy + 1.0f) + kSpan * kSpan) +
                            kSpan * fabsf(r.Bq) * (d.hx + (float)kTileW);
        r.h = (swing < 64.0f) ? exp2f(8.0f * r.Cq) : -1.0f;  // NaN swing compares false -> steep
        const float4 *rv = reinterpret_cast<const float4 *>(&r);
        rec[row * 3 + 0] = rv[0];
        rec[row * 3 + 1] = rv[1];
        rec[row * 3 + 2] = rv[2];
        // A splat with alpha == 0 leaves every pixel unchanged ((1-0)*C + 0*col == C):
        // give it an empty cull box so no tile ever lists it.
        uint2 box;
        const bool visible = d.a > 0.0f && d.x1 >= d.x0 && d.y1 >= d.y0;
        box.x = visible ? (uint32_t)r.xpack : 0xffff7fffu;  // x0 = 32767, x1 = -1
        box.y = visible ? (uint32_t)r.ypack : 0xffff7fffu;
        aabb[row] = box;
    }
}

__global__ void __launch_bounds__(kDecodeThreads)
encode_kernel(const float *__restrict__ axes, int cols, int64_t rows, float *__restrict__ chol)
{
    extern __shared__ __align__(16) float sm[];
    pdl_wait();
    pdl_trigger();
    float tmp[9];
    const float *g = stage_rows(axes, cols, rows, sm, tmp);
    if (g == nullptr) return;
    const int64_t row = (int64_t)blockIdx.x * kDecodeThreads + threadIdx.x;
    const Chol c = encode_axes(g);
    float *o = chol + row * 9;
    o[0] = c.x;
    o[1] = c.y;
    o[2] = c.log_l11;
    o[3] = c.log_l22;
    o[4] = c.l21;
    o[5] = c.r;
    o[6] = c.g;
    o[7] = c.b;
    o[8] = c.a;
}

inline size_t stage_bytes(int cols)
{
    return cols <= kDecodeStageMaxCols ? (size_t)kDecodeThreads * cols * sizeof(float) : 16;
}

}  // namespace

cudaError_t launch_decode(const float *d_genomes, int layout, int64_t rows, int cols, int H, int W,
                          float k_sigma, float4 *rec, uint2 *aabb, float *raw_f, int32_t *raw_i,
                          int *counters, int n_counters, cudaStream_t stream)
{
    if (rows <= 0) {
        // no splats: the raster still needs its per-candidate counters cleared
        if (counters != nullptr && n_counters > 0)
            return cudaMemsetAsync(counters, 0, (size_t)n_counters * sizeof(int), stream);
        return cudaSuccess;
    }
    const unsigned grid = (unsigned)((rows + kDecodeThreads - 1) / kDecodeThreads);
    const size_t smem = stage_bytes(cols);
    if (layout == GGS_LAYOUT_AXES_ANGLE)
        return launch_kernel(decode_kernel<true>, grid, kDecodeThreads, smem, stream, d_genomes, cols,
                             rows, H, W, k_sigma, rec, aabb, raw_f, raw_i, counters, n_counters);
    return launch_kernel(decode_kernel<false>, grid, kDecodeThreads, smem, stream, d_genomes, cols,
                         rows, H, W, k_sigma, rec, aabb, raw_f, raw_i, counters, n_counters);
}

cudaError_t launch_encode(const float *d_axes, int64_t rows, int cols, float *d_chol,
                          cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((rows + kDecodeThreads - 1) / kDecodeThreads);
    return launch_kernel(encode_kernel, grid, kDecodeThreads, stage_bytes(cols), stream, d_axes, cols,
                         rows, d_chol);
}

}  // namespace ggs
