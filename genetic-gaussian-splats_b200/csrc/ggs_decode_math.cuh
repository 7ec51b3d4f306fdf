// Per-splat decode arithmetic, shared by the decode kernel (ggs_decode.cu) and by the raster's
// fused-decode variant (ggs_raster.cu).
//
// Restates, per splat and on device, the reference's
//   axes_angle_to_cholesky / genome_to_renderer[_batched]   modules/encode.py:5-24, 28-59, 63-79
//   _preprocess_genome                                     modules/render.py:9-47
// The integer AABB is semantics (splats are hard-clipped to it, render.py:175-177), so this
// file reproduces the reference's fp32 operation order exactly: every arithmetic step is an
// explicitly rounded intrinsic (__fmul_rn/__fadd_rn/... are never contracted into FMAs),
// transcendentals are the IEEE-accurate libdevice expf/logf/sinf/cosf (the functions torch's
// CUDA kernels call), sqrt and division are correctly rounded.  Every translation unit that
// includes this header is compiled with -fmad=false as a second line of defence (and so that
// the inlined libdevice code is the same in all of them: the fused and the two-kernel paths
// must produce the same bits).
#pragma once

#include <math.h>

#include "ggs_common.cuh"

namespace ggs {

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

// Decoded splat -> the raster's 48-byte record and packed cull box (ggs_common.cuh).
//   f = exp(-0.5*quad)*a  ==  2^(A qx^2 + Bq qx qy + Cq qy^2 + log2 a)
__device__ __forceinline__ void make_record(const Decoded &d, SplatRec &r, uint2 &box)
{
    const float kHalfLog2e = 0.72134752044448170368f;
    const float kLog2e = 1.44269504088896340736f;
    r.cx = d.cx;
    r.cy = d.cy;
    r.A = __fmul_rn(-kHalfLog2e, d.sxx);
    r.Bq = __fmul_rn(-kLog2e, d.sxy);
    r.Cq = __fmul_rn(-kHalfLog2e, d.syy);
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
    const float reach_y = __fadd_rn(__fmul_rn(2.0f * kSpan, __fadd_rn(d.hy, 1.0f)), kSpan * kSpan);
    const float reach_x = __fmul_rn(kSpan, __fadd_rn(d.hx, (float)kTileW));
    const float swing = __fadd_rn(__fmul_rn(fabsf(r.Cq), reach_y), __fmul_rn(fabsf(r.Bq), reach_x));
    r.h = (swing < 64.0f) ? exp2f(__fmul_rn(8.0f, r.Cq)) : -1.0f;  // NaN swing compares false -> steep
    // A splat with alpha == 0 leaves every pixel unchanged ((1-0)*C + 0*col == C):
    // give it an empty cull box so no tile ever lists it.
    const bool visible = d.a > 0.0f && d.x1 >= d.x0 && d.y1 >= d.y0;
    box.x = visible ? (uint32_t)r.xpack : 0xffff7fffu;  // x0 = 32767, x1 = -1
    box.y = visible ? (uint32_t)r.ypack : 0xffff7fffu;
}

}  // namespace ggs
