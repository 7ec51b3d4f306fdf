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
//   * per pixel the work is FADD + 2 FFMA (exponent, Horner in qy) + MUFU.EX2 + 3 x (FADD+FFMA).
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

struct Pixels {
    float r[kRowsPerThread], g[kRowsPerThread], b[kRowsPerThread];
};

// Blend splat `s` of the staged list into this thread's 8 pixels (render.py:175-196).
#define GGS_PIXEL(i)                                             \
    {                                                            \
        const float qy = dy + (float)(i);                        \
        const float e = fmaf(fmaf(Cq, qy, t1), qy, t0);          \
        const float f = ex2_approx(e);                           \
        px.r[i] = fmaf(f, cr - px.r[i], px.r[i]);                \
        px.g[i] = fmaf(f, cg - px.g[i], px.g[i]);                \
        px.b[i] = fmaf(f, cb - px.b[i], px.b[i]);                \
    }

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
        if (lo == 0 && hi == kRowsPerThread - 1) {
#pragma unroll
            for (int i = 0; i < kRowsPerThread; ++i) GGS_PIXEL(i)
        } else {
#pragma unroll
            for (int i = 0; i < kRowsPerThread; ++i)
                if (i >= lo && i <= hi) GGS_PIXEL(i)
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
    for (int i = 0; i < kRowsPerThread; ++i) {
        px.r[i] = bg_r;  // render.py:236-237
        px.g[i] = bg_g;
        px.b[i] = bg_b;
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
        if (X < W && Y < H) {
            const float cr = clamp01(px.r[i]), cg = clamp01(px.g[i]), cb = clamp01(px.b[i]);
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
