// Importance (edge / detail) weight mask on the device: the input `weight_mask` of the fitness
// (reference: modules/mask.py:29-83, compute_importance_mask).
//
// Computed once per run on a small image, so the kernels are simple per-pixel kernels over
// HBM-resident planes (a 256x256 run moves < 10 MB in all): what matters is that the arithmetic
// follows the reference step by step -- torch's bilinear resize (align_corners=False), strided
// average pooling, zero-padded 3x3 Sobel and box sums with count_include_pad, torch.quantile's
// linear interpolation between exact order statistics, pow -- so the mask, and through it the
// masked fitness, matches.  This TU is compiled with -fmad=false (one rounding per operation,
// as in the reference's separate torch ops).
#include <math.h>

#include "ggs_common.cuh"

namespace ggs {
namespace {

constexpr int kPix = 256;      // threads per CTA of the per-pixel kernels
constexpr int kSelect = 1024;  // threads of the single-CTA order-statistic kernel

inline int grid_for(int64_t n) { return (int)((n + kPix - 1) / kPix); }

// Source coordinate of torch's upsample_bilinear2d, align_corners=False:
//   src = scale*(dst + 0.5) - 0.5, clamped at 0;  i1 = i0 + (i0 < in-1);  l1 = src - i0.
struct Tap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Tap bilinear_tap(int dst, float scale, int in)
{
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.0f ? 0.0f : src;
    Tap t;
    t.i0 = min((int)src, in - 1);
    t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0);
    t.l1 = src - (float)t.i0;
    t.l0 = 1.0f - t.l1;
    return t;
}

// mask.py:45-48 + :6-10: optional /255, bilinear resize of the three channels to the work size,
// Rec.709 luma.
__global__ void __launch_bounds__(kPix)
luma_kernel(const float *__restrict__ img, int H0, int W0, int H, int W, int div255,
            float *__restrict__ y)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= (int64_t)H * W) return;
    const int oy = (int)(i / W), ox = (int)(i - (int64_t)oy * W);
    const Tap ty = bilinear_tap(oy, (float)H0 / (float)H, H0);
    const Tap tx = bilinear_tap(ox, (float)W0 / (float)W, W0);
    float c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p00 = __ldg(img + ((int64_t)ty.i0 * W0 + tx.i0) * 3 + k);
        float p01 = __ldg(img + ((int64_t)ty.i0 * W0 + tx.i1) * 3 + k);
        float p10 = __ldg(img + ((int64_t)ty.i1 * W0 + tx.i0) * 3 + k);
        float p11 = __ldg(img + ((int64_t)ty.i1 * W0 + tx.i1) * 3 + k);
        if (div255) {
            p00 = p00 / 255.0f;
            p01 = p01 / 255.0f;
            p10 = p10 / 255.0f;
            p11 = p11 / 255.0f;
        }
        c[k] = ty.l0 * (tx.l0 * p00 + tx.l1 * p01) + ty.l1 * (tx.l0 * p10 + tx.l1 * p11);
    }
    y[i] = 0.2126f * c[0] + 0.7152f * c[1] + 0.0722f * c[2];
}

// F.avg_pool2d(y, kernel_size=s, stride=s) (mask.py:54): [H,W] -> [H/s, W/s].
__global__ void __launch_bounds__(kPix)
pool_kernel(const float *__restrict__ in, int W, int s, int h, int w, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= (int64_t)h * w) return;
    const int oy = (int)(i / w), ox = (int)(i - (int64_t)oy * w);
    float sum = 0.0f;
    for (int a = 0; a < s; ++a)
        for (int b = 0; b < s; ++b) sum += __ldg(in + (int64_t)(oy * s + a) * W + ox * s + b);
    out[i] = sum / (float)(s * s);
}

// mask.py:13-18: 3x3 Sobel taps (cross-correlation, zero padding), sqrt(gx^2 + gy^2 + 1e-12).
// add != 0: out += e (the full-resolution scale, mask.py:58-59).
__global__ void __launch_bounds__(kPix)
sobel_kernel(const float *__restrict__ in, int h, int w, float *__restrict__ out, int add)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= (int64_t)h * w) return;
    const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
    float v[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int yy = y + a - 1, xx = x + b - 1;
            v[a][b] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(in + (int64_t)yy * w + xx) : 0.0f;
        }
    const float gx = (v[0][2] - v[0][0]) + 2.0f * (v[1][2] - v[1][0]) + (v[2][2] - v[2][0]);
    const float gy = (v[2][0] - v[0][0]) + 2.0f * (v[2][1] - v[0][1]) + (v[2][2] - v[0][2]);
    const float e = sqrtf(gx * gx + gy * gy + 1e-12f);
    out[i] = add ? out[i] + e : e;
}

// mask.py:56-59: bilinear upsample of a coarse edge map to the work size, added to `edges`.
__global__ void __launch_bounds__(kPix)
upsample_add_kernel(const float *__restrict__ in, int h, int w, int H, int W,
                    float *__restrict__ edges)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= (int64_t)H * W) return;
    const int oy = (int)(i / W), ox = (int)(i - (int64_t)oy * W);
    const Tap ty = bilinear_tap(oy, (float)h / (float)H, h);
    const Tap tx = bilinear_tap(ox, (float)w / (float)W, w);
    const float p00 = __ldg(in + (int64_t)ty.i0 * w + tx.i0), p01 = __ldg(in + (int64_t)ty.i0 * w + tx.i1);
    const float p10 = __ldg(in + (int64_t)ty.i1 * w + tx.i0), p11 = __ldg(in + (int64_t)ty.i1 * w + tx.i1);
    const float e = ty.l0 * (tx.l0 * p00 + tx.l1 * p01) + ty.l1 * (tx.l0 * p10 + tx.l1 * p11);
    edges[i] = edges[i] + e;
}

// F.avg_pool2d(x, k, stride=1, padding=k/2) with count_include_pad (divide by k*k everywhere).
//   mode 0: out = box(in)                                    (smoothing, mask.py:76)
//   mode 1: out = max(box(in^2) - box(in)^2, 0)              (local variance, mask.py:21-25)
__global__ void __launch_bounds__(kPix)
box_kernel(const float *__restrict__ in, int H, int W, int k, int mode, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= (int64_t)H * W) return;
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    const int half = k / 2;
    float s1 = 0.0f, s2 = 0.0f;
    for (int a = -half; a <= half; ++a) {
        const int yy = y + a;
        if (yy < 0 || yy >= H) continue;
        for (int b = -half; b <= half; ++b) {
            const int xx = x + b;
            if (xx < 0 || xx >= W) continue;
            const float v = __ldg(in + (int64_t)yy * W + xx);
            s1 += v;
            s2 += v * v;
        }
    }
    const float area = (float)(k * k);
    const float m1 = s1 / area;
    if (mode == 0) {
        out[i] = m1;
    } else {
        const float m2 = s2 / area;
        out[i] = fmaxf(m2 - m1 * m1, 0.0f);
    }
}

// ---- torch.quantile(t.flatten(), q), linear interpolation (mask.py:66-68) -------------------
// Exact order statistics by radix select: four passes over the data, 8 key bits per pass, one
// shared-memory histogram per wanted rank (4 ranks: floor and ceil positions of the two
// quantiles).  One CTA: the data is a single image plane and this runs a handful of times per
// run.  out2 = (q_lo, q_hi).
__device__ __forceinline__ unsigned sort_key(float v)
{
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ float torch_lerp(float a, float b, float w)
{
    const float d = b - a;
    return (w < 0.5f) ? a + w * d : b - d * (1.0f - w);
}

__global__ void __launch_bounds__(kSelect)
quantile_kernel(const float *__restrict__ v, int64_t n, float rank_lo, float rank_hi,
                float *__restrict__ out2)
{
    __shared__ unsigned hist[4][256];
    __shared__ unsigned prefix[4];
    __shared__ unsigned long long want[4];
    const int tid = threadIdx.x;
    if (tid < 4) {
        const float r = (tid < 2) ? rank_lo : rank_hi;
        want[tid] = (unsigned long long)((tid & 1) ? ceilf(r) : floorf(r));
        prefix[tid] = 0u;
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned known = pass ? (0xffffffffu << (shift + 8)) : 0u;
        for (int i = tid; i < 4 * 256; i += kSelect) (&hist[0][0])[i] = 0u;
        __syncthreads();
        unsigned pre[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) pre[t] = prefix[t];
        const int64_t n_round = (n + kSelect - 1) / kSelect * kSelect;  // whole warps stay converged
        for (int64_t i = tid; i < n_round; i += kSelect) {
            const bool live = i < n;
            const unsigned key = live ? sort_key(__ldg(v + i)) : 0u;
            const unsigned digit = (key >> shift) & 255u;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool mine = live && ((key & known) == pre[t]);
                // one shared-memory atomic per distinct digit in the warp
                const unsigned peers = __match_any_sync(0xffffffffu, mine ? digit : 256u);
                if (mine && (__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[t][digit], __popc(peers));
            }
        }
        __syncthreads();
        if (tid < 4) {
            unsigned long long k = want[tid];
            unsigned d = 0;
            for (; d < 255u; ++d) {
                const unsigned c = hist[tid][d];
                if (k < c) break;
                k -= c;
            }
            want[tid] = k;
            prefix[tid] |= d << shift;
        }
        __syncthreads();
    }
    if (tid < 2) {
        const float r = tid ? rank_hi : rank_lo;
        const float below = key_value(prefix[2 * tid]), above = key_value(prefix[2 * tid + 1]);
        out2[tid] = torch_lerp(below, above, r - floorf(r));
    }
}

// mask.py:66-69: ((t - ql) / (qh - ql + 1e-12)).clamp(0, 1), in place.
__global__ void __launch_bounds__(kPix)
norm_kernel(float *__restrict__ t, int64_t n, const float *__restrict__ q2)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= n) return;
    const float ql = q2[0], qh = q2[1];
    const float r = (t[i] - ql) / ((qh - ql) + 1e-12f);
    t[i] = fminf(fmaxf(r, 0.0f), 1.0f);
}

// mask.py:73: w_edge*E + w_var*V.
__global__ void __launch_bounds__(kPix)
mix_kernel(const float *__restrict__ e, const float *__restrict__ v, int64_t n, float w_edge,
           float w_var, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i < n) out[i] = w_edge * e[i] + w_var * v[i];
}

// mask.py:79-86: pow(gamma), lift to `floor`, blend towards 1 by (1 - strength).
__global__ void __launch_bounds__(kPix)
shape_kernel(const float *__restrict__ m, int64_t n, float gamma, float one_minus_floor,
             float floor_, float one_minus_strength, float strength, int blend,
             float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * kPix + threadIdx.x;
    if (i >= n) return;
    float v = powf(m[i], gamma);
    v = one_minus_floor * v + floor_;
    if (blend) v = one_minus_strength * 1.0f + strength * v;
    out[i] = v;
}

}  // namespace

size_t mask_workspace_bytes(int H, int W)
{
    // y, edges, var, scratch (pooled + its Sobel), quantile results
    return (size_t)4 * align_up((size_t)H * W * sizeof(float), 256) + 256;
}

cudaError_t launch_importance_mask(const float *d_image, int H0, int W0, int H, int W, int div255,
                                   const int *scales, int n_scales, float w_edge, float w_var,
                                   float gamma, float one_minus_floor, float floor_,
                                   int smooth, float one_minus_strength, float strength,
                                   int blend, float *d_mask, void *d_ws, cudaStream_t st)
{
    const int64_t n = (int64_t)H * W;
    const size_t plane = align_up((size_t)n * sizeof(float), 256);
    char *base = static_cast<char *>(d_ws);
    float *y = reinterpret_cast<float *>(base);
    float *edges = reinterpret_cast<float *>(base + plane);
    float *var = reinterpret_cast<float *>(base + 2 * plane);
    float *scratch = reinterpret_cast<float *>(base + 3 * plane);
    float *q2 = reinterpret_cast<float *>(base + 4 * plane);
    const int g = grid_for(n);
    // torch.quantile: ranks = q * (n - 1), held in the input's dtype (float32)
    const float rank_lo = 0.02f * (float)(n - 1), rank_hi = 0.98f * (float)(n - 1);
    auto normalise = [&](float *t) {
        quantile_kernel<<<1, kSelect, 0, st>>>(t, n, rank_lo, rank_hi, q2);
        norm_kernel<<<g, kPix, 0, st>>>(t, n, q2);
    };

    luma_kernel<<<g, kPix, 0, st>>>(d_image, H0, W0, H, W, div255, y);
    cudaError_t e = cudaMemsetAsync(edges, 0, (size_t)n * sizeof(float), st);  // mask.py:51
    if (e != cudaSuccess) return e;
    for (int k = 0; k < n_scales; ++k) {
        const int s = scales[k];
        if (s > 1) {
            const int h = H / s, w = W / s;
            float *pooled = scratch, *sob = scratch + (size_t)h * w;
            pool_kernel<<<grid_for((int64_t)h * w), kPix, 0, st>>>(y, W, s, h, w, pooled);
            sobel_kernel<<<grid_for((int64_t)h * w), kPix, 0, st>>>(pooled, h, w, sob, 0);
            upsample_add_kernel<<<g, kPix, 0, st>>>(sob, h, w, H, W, edges);
        } else {
            sobel_kernel<<<g, kPix, 0, st>>>(y, H, W, edges, 1);
        }
    }
    box_kernel<<<g, kPix, 0, st>>>(y, H, W, 9, 1, var);  // mask.py:62
    normalise(edges);
    normalise(var);
    float *m = y;  // the luma plane is no longer needed
    mix_kernel<<<g, kPix, 0, st>>>(edges, var, n, w_edge, w_var, m);
    normalise(m);
    if (smooth > 0) {
        box_kernel<<<g, kPix, 0, st>>>(m, H, W, smooth, 0, edges);
        m = edges;
        normalise(m);
    }
    shape_kernel<<<g, kPix, 0, st>>>(m, n, gamma, one_minus_floor, floor_, one_minus_strength,
                                     strength, blend, d_mask);
    return cudaGetLastError();
}

}  // namespace ggs
