// Fused tile rasteriser + fitness reduction: one launch for the whole population.
//
// Replaces, for every candidate at once,
//   _gpu_bin_splats_to_tiles      modules/render.py:51-118   (per-tile ordered splat lists)
//   _render_tile_over_kernel      modules/render.py:121-200  (falloff + "over" blend)
//   canvas fill / final clamp     modules/render.py:236-237, :252
//   squared error + reductions    modules/fitness.py:16-31
//
// One CTA = one (candidate, 32x32 tile).  The CTA streams the candidate's packed AABBs in
// genome order, 128 per round; a ballot/popcount prefix compacts the splats that touch the tile
// *in order* into a shared-memory list of 48-byte records (no global sort, no host sync), and
// whenever the list fills (or the genome ends) the four warps composite it.  Warp w owns the
// 32x8 band of rows [8w, 8w+8): lane = pixel column, 8 vertically adjacent pixels per thread.
// With that mapping
//   * the AABB row test is warp-uniform: rows outside [y0,y1] are skipped by uniform branches,
//   * the AABB column test is one select per (thread, splat) that sets the exponent to -inf,
//   * the falloff along a thread's pixel column follows a multiplicative recurrence, and the
//     blends run as packed FADD2/FFMA2 on row pairs (see composite_list).
// Colours stay in registers from the first splat to the fitness reduction; images are written
// only when asked for.  Per-tile partial sums are combined in a fixed order by the last CTA of
// each candidate, so fitness is bit-reproducible.
#include "ggs_common.cuh"

namespace ggs {
namespace {

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: one issue slot, two lanes of work) ----
typedef unsigned long long f2_t;

__device__ __forceinline__ f2_t pack2(float lo, float hi)
{
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2_t bcast2(float v) { return pack2(v, v); }
__device__ __forceinline__ void unpack2(f2_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c)
{
    f2_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b)
{
    f2_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b)
{
    f2_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Pixel state of one thread: 8 vertically adjacent pixels as 4 packed row pairs (2k, 2k+1).
struct Pixels {
    f2_t r[kRowsPerThread / 2], g[kRowsPerThread / 2], b[kRowsPerThread / 2];
};

// Exact per-pixel blend (render.py:189-196): exponent by Horner in qy, one MUFU.EX2 per pixel.
#define GGS_BLEND1(cr_, cg_, cb_, i)                          \
    {                                                         \
        const float qy = dy + (float)(i);                     \
        const float e = fmaf(fmaf(Cq, qy, t1), qy, t0);       \
        const float f = ex2_approx(e);                        \
        cr_ = fmaf(f, cr - cr_, cr_);                         \
        cg_ = fmaf(f, cg - cg_, cg_);                         \
        cb_ = fmaf(f, cb - cb_, cb_);                         \
    }

// Blend the staged list into this thread's pixels, in list (= genome) order.
//
// Fast path (splat covers all 8 rows of the band, exponent varies gently): the Gaussian along
// the thread's pixel column is the exponential of a quadratic, so with stride-2 steps
//     f(i+2) = f(i) * g(i),   g(i+2) = g(i) * h,   h = 2^(8*Cq)
// the 8 falloffs come from 4 MUFU.EX2 (f0, f1, g0, g1) and packed multiplies, and the three
// colour blends are FADD2 + FFMA2 on row pairs: 8 issue slots per 2 pixels instead of 20.
// Slow path (partial band, or a "steep" splat whose exponent changes too fast for the
// recurrence to stay accurate): the exact per-pixel form, rows selected by uniform branches.
__device__ __forceinline__ void composite_list(const float4 *__restrict__ list, int cnt, int X,
                                               float Xf, int Yb, float Ybf, Pixels &px)
{
    for (int s = 0; s < cnt; ++s) {
        const float4 q2 = list[3 * s + 2];
        const int yp = __float_as_int(q2.z);
        const int y0 = (int)(short)(yp & 0xffff), y1 = yp >> 16;
        const int lo = max(y0 - Yb, 0), hi = min(y1 - Yb, kRowsPerThread - 1);
        if (lo > hi) continue;  // splat misses this warp's band (warp-uniform)
        const float4 q0 = list[3 * s + 0];
        const float4 q1 = list[3 * s + 1];
        const int xp = __float_as_int(q2.y);
        const int x0 = (int)(short)(xp & 0xffff), x1 = xp >> 16;
        const bool in_x = (X >= x0) & (X <= x1);
        const float qx = Xf - q0.x;
        const float t1 = q0.w * qx;                       // Bq*qx
        float t0 = fmaf(q0.z * qx, qx, q1.y);             // A*qx^2 + log2(alpha)
        t0 = in_x ? t0 : -INFINITY;                       // outside [x0,x1]: f = 2^-inf = 0
        const float dy = Ybf - q0.y;
        const float Cq = q1.x, cr = q1.z, cg = q1.w, cb = q2.x;
        const float h = q2.w;                             // 2^(8*Cq), or < 0 for a steep splat
        if (lo == 0 && hi == kRowsPerThread - 1 && h >= 0.0f) {
            const f2_t QY = pack2(dy, dy + 1.0f);
            const f2_t E = fma2(fma2(bcast2(Cq), QY, bcast2(t1)), QY, bcast2(t0));
            const float c4 = 4.0f * Cq;
            const f2_t D = fma2(bcast2(c4), QY, bcast2(fmaf(2.0f, t1, c4)));  // e(i+2) - e(i)
            float e0, e1, d0, d1;
            unpack2(E, e0, e1);
            unpack2(D, d0, d1);
            f2_t F = pack2(ex2_approx(e0), ex2_approx(e1));
            f2_t G = pack2(ex2_approx(d0), ex2_approx(d1));
            const f2_t H2 = bcast2(h), R2 = bcast2(cr), G2 = bcast2(cg), B2 = bcast2(cb);
#pragma unroll
            for (int k = 0; k < kRowsPerThread / 2; ++k) {
                px.r[k] = fma2(F, sub2(R2, px.r[k]), px.r[k]);
                px.g[k] = fma2(F, sub2(G2, px.g[k]), px.g[k]);
                px.b[k] = fma2(F, sub2(B2, px.b[k]), px.b[k]);
                if (k + 1 < kRowsPerThread / 2) {
                    F = mul2(F, G);
                    G = mul2(G, H2);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < kRowsPerThread / 2; ++k) {
                if (2 * k + 1 >= lo && 2 * k <= hi) {
                    float r0, r1, g0, g1, b0, b1;
                    unpack2(px.r[k], r0, r1);
                    unpack2(px.g[k], g0, g1);
                    unpack2(px.b[k], b0, b1);
                    if (2 * k >= lo) GGS_BLEND1(r0, g0, b0, 2 * k)
                    if (2 * k + 1 <= hi) GGS_BLEND1(r1, g1, b1, 2 * k + 1)
                    px.r[k] = pack2(r0, r1);
                    px.g[k] = pack2(g0, g1);
                    px.b[k] = pack2(b0, b1);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
raster_kernel(const float4 *__restrict__ rec, const uint2 *__restrict__ aabb, int N, int H, int W,
              int ntx, int ntiles, float bg_r, float bg_g, float bg_b,
              const float *__restrict__ target, const float *__restrict__ mask, int mode,
              float beta, float *__restrict__ images, float2 *__restrict__ partial,
              int *__restrict__ counter, float *__restrict__ fitness)
{
    __shared__ float4 s_list[kListCap * 3];
    __shared__ int s_wcnt[kWarps];
    __shared__ float s_red[2 * kWarps];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / ntiles;
    const int t = blockIdx.x - b * ntiles;
    const int ty = t / ntx, tx = t - ty * ntx;
    const int X0 = tx * kTileW, Y0 = ty * kTileH;
    const int X1 = X0 + kTileW - 1, Y1 = Y0 + kTileH - 1;
    const int X = X0 + lane, Yb = Y0 + warp * kRowsPerThread;
    const float Xf = (float)X, Ybf = (float)Yb;

    Pixels px;
#pragma unroll
    for (int k = 0; k < kRowsPerThread / 2; ++k) {
        px.r[k] = bcast2(bg_r);  // render.py:236-237
        px.g[k] = bcast2(bg_g);
        px.b[k] = bcast2(bg_b);
    }

    const float4 *recb = rec + (int64_t)b * N * 3;
    const uint2 *boxb = aabb + (int64_t)b * N;

    int cnt = 0;
    for (int base = 0; base < N; base += kThreads) {
        const int i = base + tid;
        bool hit = false;
        if (i < N) {
            const uint2 box = __ldg(boxb + i);
            const int x0 = (int)(short)(box.x & 0xffff), x1 = (int)box.x >> 16;
            const int y0 = (int)(short)(box.y & 0xffff), y1 = (int)box.y >> 16;
            hit = (x1 >= X0) & (x0 <= X1) & (y1 >= Y0) & (y0 <= Y1);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int c = s_wcnt[w];
            pre += (w < warp) ? c : 0;
            tot += c;
        }
        if (hit) {
            const int pos = cnt + pre + __popc(m & ((1u << lane) - 1u));
            const float4 *src = recb + (int64_t)i * 3;
            s_list[pos * 3 + 0] = __ldg(src + 0);
            s_list[pos * 3 + 1] = __ldg(src + 1);
            s_list[pos * 3 + 2] = __ldg(src + 2);
        }
        cnt += tot;
        __syncthreads();
        if (cnt > kListCap - kThreads || base + kThreads >= N) {
            composite_list(s_list, cnt, X, Xf, Yb, Ybf, px);
            cnt = 0;
            __syncthreads();
        }
    }

    // Epilogue: clamp (render.py:252), optional image store, squared error (fitness.py:16-31).
    float num = 0.0f, den = 0.0f;
    const bool want_fit = (target != nullptr);
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
        const int Y = Yb + i;
        float pr[2], pg[2], pb[2];
        unpack2(px.r[i >> 1], pr[0], pr[1]);
        unpack2(px.g[i >> 1], pg[0], pg[1]);
        unpack2(px.b[i >> 1], pb[0], pb[1]);
        if (X < W && Y < H) {
            const float cr = clamp01(pr[i & 1]), cg = clamp01(pg[i & 1]), cb = clamp01(pb[i & 1]);
            const int64_t p = (int64_t)Y * W + X;
            if (images != nullptr) {
                float *o = images + ((int64_t)b * H * W + p) * 3;
                o[0] = cr;
                o[1] = cg;
                o[2] = cb;
            }
            if (want_fit) {
                const float dr = cr - __ldg(target + 3 * p + 0);
                const float dg = cg - __ldg(target + 3 * p + 1);
                const float db = cb - __ldg(target + 3 * p + 2);
                float w = 1.0f;
                if (mode == GGS_MODE_MASK)
                    w = __ldg(mask + p);
                else if (mode == GGS_MODE_BOOST)
                    w = 1.0f + beta * clamp01(__ldg(mask + p));
                num += (dr * dr) * w + (dg * dg) * w + (db * db) * w;
                den += w;
            }
        }
    }
    if (!want_fit) return;

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (lane == 0) {
        s_red[warp] = num;
        s_red[kWarps + warp] = den;
    }
    __syncthreads();
    if (tid == 0) {
        float n = 0.0f, d = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            n += s_red[w];
            d += s_red[kWarps + w];
        }
        partial[blockIdx.x] = make_float2(n, d);
        __threadfence();
        const int ticket = atomicAdd(counter + b, 1);
        s_last = (ticket == ntiles - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
        // Last tile of this candidate: combine the per-tile partials in tile order.
        __threadfence();
        const volatile float2 *pb = partial + (int64_t)b * ntiles;
        double n = 0.0, d = 0.0;
        for (int k = lane; k < ntiles; k += 32) {
            n += (double)pb[k].x;
            d += (double)pb[k].y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n += __shfl_xor_sync(0xffffffffu, n, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == 0) {
            const double P = (double)H * (double)W;
            float fit;
            if (mode == GGS_MODE_PLAIN)
                fit = (float)(n / (3.0 * P));                            // fitness.py:19
            else if (mode == GGS_MODE_MASK)
                fit = (float)n / ((float)d + 1e-12f);                    // fitness.py:29-31
            else
                fit = (float)(n / (3.0 * P)) / ((float)(d / P) + 1e-12f);  // fitness.py:23-27
            fitness[b] = fit;
            counter[b] = 0;  // ready for the next launch on this workspace
        }
    }
}

}  // namespace

cudaError_t launch_raster(const Workspace &ws, int B, int N, int H, int W, const float bg[3],
                          const float *d_target, const float *d_mask, int mode, float beta,
                          float *d_fitness, float *d_images, cudaStream_t stream)
{
    if (B <= 0) return cudaSuccess;
    const int ntx = tiles_x(W), ntiles = ntx * tiles_y(H);
    const int64_t grid = (int64_t)B * ntiles;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    raster_kernel<<<(unsigned)grid, kThreads, 0, stream>>>(
        ws.rec, ws.aabb, N, H, W, ntx, ntiles, bg[0], bg[1], bg[2], d_target, d_mask, mode, beta,
        d_images, ws.partial, ws.counter, d_fitness);
    return cudaGetLastError();
}

}  // namespace ggs
