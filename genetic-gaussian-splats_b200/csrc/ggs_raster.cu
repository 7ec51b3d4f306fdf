// Fused tile rasteriser + fitness reduction: one launch for the whole population.
//
// Replaces, for every candidate at once,
//   _gpu_bin_splats_to_tiles      modules/render.py:51-118   (per-tile ordered splat lists)
//   _render_tile_over_kernel      modules/render.py:121-200  (falloff + "over" blend)
//   canvas fill / final clamp     modules/render.py:236-237, :252
//   squared error + reductions    modules/fitness.py:16-31
//
// One CTA = one (candidate, tile of 32 x kTileH pixels).  The CTA streams the candidate's packed
// AABBs, 256 per round; a ballot/popcount prefix compacts the splats that touch the tile *in
// order* into a shared-memory list of 48-byte records (no global sort, no host sync; the q0/q1
// parts are gathered with cp.async), and whenever the list fills (or the genome ends) the warps
// composite it.  Warp w owns the band of rows [R*w, R*w + R), R = kRowsPerThread: lane = pixel
// column, R vertically adjacent pixels per thread.  With that mapping
//   * the AABB row test is warp-uniform: the thread that stages a record precomputes, per band,
//     which rows it covers (one byte per warp), so a warp classifies a splat with one PRMT;
//   * the AABB column test is a per-tile lane mask precomputed the same way: one LOP3 + one
//     select per (thread, splat) sets the exponent to -inf outside [x0, x1];
//   * the falloff along a thread's pixel column follows a multiplicative recurrence and the
//     blends run as packed FMUL2/FFMA2/FADD2 on row pairs (see composite_list).
//
// Compositing order.  The reference paints splats in genome order with the "over" operator
// C <- (1-f) C + f col (render.py:194-196).  The same image is
//     C = sum_n  f_n col_n  prod_{m>n} (1 - f_m)   +   bg prod_m (1 - f_m),
// which this kernel evaluates front to back: it walks the genome from the LAST splat to the
// first, keeping per pixel the accumulated colour and the transmittance T = prod (1 - f_m).
// That form needs 5 packed operations per row pair instead of 6 and is algebraically identical;
// parity with the reference is checked to 1e-4 absolute on every pixel (tests/).
//
// Colours stay in registers from the first splat to the fitness reduction; images are written
// only when asked for.  Per-tile partial sums are combined in a fixed order by the last CTA of
// each candidate, so fitness is bit-reproducible.
//
// Code generation notes (all measured):
//   * the pixel state lives in *named PTX registers* that only in-place PTX touches; as C++
//     values ptxas renamed the 64-bit accumulators out of place in most builds and paid ~20
//     MOVs per splat (tests/test_cpu_sass.py guards this);
//   * per-thread constants used in the loop are routed through a shuffle so ptxas cannot
//     rematerialise them from %tid.x once per list entry.
#include "ggs_common.cuh"

namespace ggs {
namespace {

#ifndef GGS_MIN_BLOCKS
#define GGS_MIN_BLOCKS (GGS_ROWS == 8 ? (GGS_WARPS == 4 ? 8 : 16) : (GGS_WARPS == 4 ? 4 : 9))
#endif

constexpr int kPairs = kRowsPerThread / 2;
#ifndef GGS_SCAN_CHUNK
#define GGS_SCAN_CHUNK 256
#endif
constexpr int kScanPerThread = GGS_SCAN_CHUNK / kThreads;  // 256 records examined per round
static_assert(GGS_SCAN_CHUNK <= kListCap, "a scan round must fit the list");
constexpr int kScanChunk = kThreads * kScanPerThread;
#ifndef GGS_SAT_EVERY
#define GGS_SAT_EVERY 8
#endif
#ifndef GGS_BOX_PREFETCH
#define GGS_BOX_PREFETCH 1
#endif
#ifndef GGS_DENSE_STAGE
#define GGS_DENSE_STAGE 1
#endif
#if GGS_DENSE_STAGE && !GGS_BOX_PREFETCH
#error "GGS_DENSE_STAGE needs GGS_BOX_PREFETCH"
#endif
constexpr int kSatEvery = GGS_SAT_EVERY;                 // list entries between saturation votes
constexpr float kOpaque = 2.384185791015625e-07f;        // 2^-22: transmittance counted as zero
static_assert(kScanPerThread >= 1, "at most 256 threads per CTA");

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 16-byte asynchronous global -> shared copy (SASS LDGSTS), cached in L1: the tiles of one
// candidate run on neighbouring CTAs and re-read the same records.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b)
{
    f2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- pixel state: named PTX registers ggs_{r,g,b,t}<pair> -------------------------------------
// Pair k of a thread holds rows (2k, 2k+1) of its column: accumulated colour (premultiplied,
// front to back) and transmittance.  GGS_PAIRS(M) expands M(0) ... M(kPairs-1).
#if GGS_ROWS == 8
#define GGS_PX_DECLARE() asm volatile(".reg .b64 ggs_r<4>, ggs_g<4>, ggs_b<4>, ggs_t<4>, ggs_F, ggs_G;")
#define GGS_PAIRS(M) M(0) M(1) M(2) M(3)
#else
#define GGS_PX_DECLARE() asm volatile(".reg .b64 ggs_r<8>, ggs_g<8>, ggs_b<8>, ggs_t<8>, ggs_F, ggs_G;")
#define GGS_PAIRS(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
#endif
#define GGS_PX_INIT(k, ta_, tb_)                                                          \
    asm volatile("mov.b64 ggs_r" #k ", 0;\n\tmov.b64 ggs_g" #k ", 0;\n\tmov.b64 ggs_b" #k   \
                 ", 0;\n\tmov.b64 ggs_t" #k ", {%0, %1};" ::"f"(ta_), "f"(tb_));
// The falloff pair of the current rows (ggs_F) and its stride-2 ratio (ggs_G) are named too.
#define GGS_SET_F(f0_, f1_) asm volatile("mov.b64 ggs_F, {%0, %1};" ::"f"(f0_), "f"(f1_));
#define GGS_SET_G(g0_, g1_) asm volatile("mov.b64 ggs_G, {%0, %1};" ::"f"(g0_), "f"(g1_));
#define GGS_STEP_F() asm volatile("mul.rn.f32x2 ggs_F, ggs_F, ggs_G;");
#define GGS_STEP_G(H_) asm volatile("mul.rn.f32x2 ggs_G, ggs_G, %0;" ::"l"(H_));
// One row pair, front to back (render.py:194-196 rearranged): W = F*T, C += W*col, T -= W.
#define GGS_PX_BLEND(k)                                                                   \
    asm volatile("{\n\t.reg .b64 w;\n\t"                                                 \
                 "mul.rn.f32x2 w, ggs_F, ggs_t" #k ";\n\t"                                \
                 "fma.rn.f32x2 ggs_r" #k ", w, %0, ggs_r" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_g" #k ", w, %1, ggs_g" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_b" #k ", w, %2, ggs_b" #k ";\n\t"                      \
                 "sub.rn.f32x2 ggs_t" #k ", ggs_t" #k ", w;\n\t}" ::"l"(R2), "l"(G2),    \
                 "l"(B2));
#define GGS_PX_READ_T(k, a_, b_) asm volatile("mov.b64 {%0, %1}, ggs_t" #k ";" : "=f"(a_), "=f"(b_));
#define GGS_PX_READ(k, r_, g_, b_, t_)                                                    \
    asm volatile("mov.b64 {%0, %1}, ggs_r" #k ";\n\tmov.b64 {%2, %3}, ggs_g" #k            \
                 ";\n\tmov.b64 {%4, %5}, ggs_b" #k ";\n\tmov.b64 {%6, %7}, ggs_t" #k ";"   \
                 : "=f"(r_[0]), "=f"(r_[1]), "=f"(g_[0]), "=f"(g_[1]), "=f"(b_[0]),       \
                   "=f"(b_[1]), "=f"(t_[0]), "=f"(t_[1]));

// Staged record (shared memory, 3 x float4), specialised for the tile by the staging thread:
//   q0 = cx, cy, A, Bq      q1 = Cq, la, r, g      q2 = b, lane mask, row code, h
// lane mask: bit l set iff column X0+l lies in [x0, x1].
// row code : byte w describes band w: 0x0f = not touched, else lo | hi << 4 (rows lo..hi of the
//            band are inside [y0, y1]), bit 7 set for a steep splat (h < 0); (R-1) << 4 = all
//            rows of a gentle splat, the only value that takes the recurrence path.
constexpr unsigned kBandMiss = 0x0fu;
constexpr unsigned kBandFull = (unsigned)(kRowsPerThread - 1) << 4;

// With 8 rows per thread hi needs 3 bits, and bit 7 of a touched band's byte flags a steep splat:
// "byte == kBandFull" is then the whole test for the recurrence path.
constexpr unsigned kSteepBit = (kRowsPerThread == 8) ? 0x80u : 0u;

__device__ __forceinline__ unsigned row_code(int y0, int y1, int Y0, bool steep)
{
    unsigned code = 0;
    const unsigned flag = steep ? kSteepBit : 0u;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const int yb = Y0 + w * kRowsPerThread;
        const int lo = max(y0 - yb, 0), hi = min(y1 - yb, kRowsPerThread - 1);
        const unsigned c = (lo > hi) ? kBandMiss : ((unsigned)(lo | (hi << 4)) | flag);
        code |= c << (8 * w);
    }
    return code;
}

__device__ __forceinline__ unsigned lane_mask(int x0, int x1, int X0)
{
    const int l0 = max(x0 - X0, 0), l1 = min(x1 - X0, kTileW - 1);
    return (0xffffffffu >> (31 - l1)) & (0xffffffffu << l0);
}

// Blend the staged list (reverse genome order) into this thread's pixels.  The exponent of the
// falloff on row i of the thread's column is e(i) = (Cq*qy + t1)*qy + t0 with qy = dy + i
// (render.py:189-192 with the constants folded at decode time), f = 2^e.
//
// Recurrence path (splat covers all rows of the band, exponent varies gently): along the
// column f is the exponential of a quadratic, so with stride-2 steps
//     f(i+2) = f(i) * g(i),   g(i+2) = g(i) * h,   h = 2^(8*Cq)
// the falloffs of a whole column come from 4 MUFU.EX2 (f0, f1, g0, g1) and packed multiplies:
// 7 packed FMA-pipe operations per row pair in all.
// Exact path (partial band, or a "steep" splat whose exponent changes too fast for the
// recurrence to stay accurate): Horner + one MUFU.EX2 per pixel; rows outside [y0, y1] get
// f = 0 (an exact no-op) through warp-uniform selects, whole pairs are skipped by uniform
// branches.
//
// Saturation: once every pixel of the band has transmittance below kOpaque, nothing further
// back can change a pixel by more than kOpaque (colours are in [0,1]), so the warp stops
// (returns false).  Checked every kSatEvery list entries with one warp vote.
template <bool kStats>
__device__ __forceinline__ bool composite_list(const float4 *__restrict__ list, int cnt,
                                               unsigned lanebit, unsigned band_sel, float Xf,
                                               float Ybf, unsigned (&work)[2])
{
    for (int s = 0; s < cnt; ++s) {
        const float4 q2 = list[3 * s + 2];
        if ((s & (kSatEvery - 1)) == kSatEvery - 1) {
            float tmax = 0.0f;
#define GGS_TMAX(k)                        \
    {                                      \
        float ta, tb;                      \
        GGS_PX_READ_T(k, ta, tb)           \
        tmax = fmaxf(tmax, fmaxf(ta, tb)); \
    }
            GGS_PAIRS(GGS_TMAX)
#undef GGS_TMAX
            if (__all_sync(0xffffffffu, tmax < kOpaque)) return false;
        }
        const unsigned c = __byte_perm(__float_as_uint(q2.z), 0u, band_sel);  // this band's byte
        if (c == kBandMiss) continue;  // warp-uniform
        const float4 q0 = list[3 * s + 0];
        const float4 q1 = list[3 * s + 1];
        const bool in_x = (__float_as_uint(q2.y) & lanebit) != 0u;
        const float qx = Xf - q0.x;
        const float t1 = q0.w * qx;                       // Bq*qx
        float t0 = fmaf(q0.z * qx, qx, q1.y);             // A*qx^2 + log2(alpha)
        t0 = in_x ? t0 : -INFINITY;                       // outside [x0,x1]: f = 2^-inf = 0
        const float dy = Ybf - q0.y;
        const f2_t QY = pack2(dy, dy + 1.0f);
        const f2_t CQ2 = bcast2(q1.x), T12 = bcast2(t1), T02 = bcast2(t0);
        const f2_t R2 = bcast2(q1.z), G2 = bcast2(q1.w), B2 = bcast2(q2.x);
        if (c == kBandFull && (kSteepBit != 0u || q2.w >= 0.0f)) {
            const f2_t E = fma2(fma2(CQ2, QY, T12), QY, T02);
            const float c4 = 4.0f * q1.x;
            const f2_t D = fma2(bcast2(c4), QY, bcast2(fmaf(2.0f, t1, c4)));  // e(i+2) - e(i)
            float e0, e1, d0, d1;
            unpack2(E, e0, e1);
            unpack2(D, d0, d1);
            GGS_SET_F(ex2_approx(e0), ex2_approx(e1))
            GGS_SET_G(ex2_approx(d0), ex2_approx(d1))
            const f2_t H2 = bcast2(q2.w);
            if (kStats) work[0] += kPairs;
#define GGS_RECUR_PAIR(k)                    \
    GGS_PX_BLEND(k)                          \
    if (k + 1 < kPairs) {                    \
        GGS_STEP_F()                         \
        if (k + 2 < kPairs) GGS_STEP_G(H2)   \
    }
            GGS_PAIRS(GGS_RECUR_PAIR)
#undef GGS_RECUR_PAIR
        } else {
            const int lo = (int)(c & 15u), hi = (int)((c & ~kSteepBit) >> 4);
#define GGS_EXACT_PAIR(k)                                                          \
    if (2 * k + 1 >= lo && 2 * k <= hi) {                                          \
        const f2_t QYk = add2(QY, bcast2((float)(2 * k)));                         \
        const f2_t E = fma2(fma2(CQ2, QYk, T12), QYk, T02);                        \
        float e0, e1;                                                              \
        unpack2(E, e0, e1);                                                        \
        const float f0 = (2 * k >= lo) ? ex2_approx(e0) : 0.0f;                    \
        const float f1 = (2 * k + 1 <= hi) ? ex2_approx(e1) : 0.0f;                \
        GGS_SET_F(f0, f1)                                                          \
        if (kStats) work[1] += 1;                                                  \
        GGS_PX_BLEND(k)                                                            \
    }
            GGS_PAIRS(GGS_EXACT_PAIR)
#undef GGS_EXACT_PAIR
        }
    }
    return true;
}

template <bool kStats>
__global__ void __launch_bounds__(kThreads, GGS_MIN_BLOCKS)
raster_kernel(const float4 *__restrict__ rec, const uint2 *__restrict__ aabb, int N, int H, int W,
              int ntx, int ntiles, float bg_r, float bg_g, float bg_b,
              const float *__restrict__ target, const float *__restrict__ mask, int mode,
              float beta, float *__restrict__ images, int image_u8, float2 *__restrict__ partial,
              int *__restrict__ counter, float *__restrict__ fitness,
              unsigned long long *__restrict__ stats)
{
    unsigned work[2] = {0u, 0u};  // kStats: row pairs blended on the recurrence / exact path
    __shared__ float4 s_list[kListCap * 3];
#if GGS_DENSE_STAGE
    __shared__ int s_wcnt[2][kScanPerThread][kWarps];
    __shared__ int s_idx[kListCap];
#else
    __shared__ int s_wcnt[kScanPerThread][kWarps];
#endif
    __shared__ float s_red[2 * kWarps];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / ntiles;
    const int t = blockIdx.x - b * ntiles;
    const int ty = t / ntx, tx = t - ty * ntx;
    const int X0 = tx * kTileW, Y0 = ty * kTileH;
    const int X1 = X0 + kTileW - 1, Y1 = Y0 + kTileH - 1;
    const int X = X0 + lane, Yb = Y0 + warp * kRowsPerThread;
    // Loop constants of the composite, pinned in registers by a (no-op) shuffle.
    const float Xf = __shfl_sync(0xffffffffu, (float)X, lane);
    const float Ybf = __shfl_sync(0xffffffffu, (float)Yb, lane);
    const unsigned lanebit = __shfl_sync(0xffffffffu, 1u << lane, lane);
    const unsigned band_sel = __shfl_sync(0xffffffffu, 0x4440u + (unsigned)warp, lane);  // PRMT: byte `warp`

    // Transmittance starts at 1 inside the image and at 0 outside it: pixels beyond the image
    // edge then take no colour and never keep a band from saturating.
    GGS_PX_DECLARE();
#define GGS_T_INIT(k)                                              \
    GGS_PX_INIT(k, (X < W && Yb + 2 * k < H) ? 1.0f : 0.0f,        \
                (X < W && Yb + 2 * k + 1 < H) ? 1.0f : 0.0f)
    GGS_PAIRS(GGS_T_INIT)
#undef GGS_T_INIT
    bool live = true;  // warp-uniform: this band still has a non-opaque pixel

    const float4 *recb = rec + (int64_t)b * N * 3;
    const uint2 *boxb = aabb + (int64_t)b * N;
    pdl_wait();  // the decode launch ahead of us has completed; nothing above reads memory
    pdl_trigger();

    // Walk the genome from its last splat to its first (front to back).  Slot j of a round
    // maps thread `tid` to record  top - 1 - (j*kThreads + tid): ascending (j, tid) is
    // descending genome order, so the ordinary ballot compaction yields the order we need.
    int cnt = 0;
#if GGS_BOX_PREFETCH
    uint2 nbox[kScanPerThread];  // the round's boxes, loaded one round ahead
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) {
        const int i0 = N - 1 - (j * kThreads + tid);
        nbox[j] = (i0 >= 0) ? __ldg(boxb + i0) : make_uint2(0xffff7fffu, 0xffff7fffu);
    }
#endif
#if GGS_DENSE_STAGE
    // A round only compacts the INDICES of the splats that touch the tile (ordered, by ballot and
    // a cross-warp prefix); the records are staged when the list is about to be composited, one
    // list entry per thread, so the staging code runs ceil(cnt / kThreads) times per flush instead
    // of once per (round, slot) with a few lanes active.
    int par = 0;
    for (int top = N; top > 0; top -= kScanChunk) {
        bool hit[kScanPerThread];
        unsigned bal[kScanPerThread];
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            const uint2 box = nbox[j];
            const int bx0 = (int)(short)(box.x & 0xffff), bx1 = (int)box.x >> 16;
            const int by0 = (int)(short)(box.y & 0xffff), by1 = (int)box.y >> 16;
            hit[j] = (bx1 >= X0) & (bx0 <= X1) & (by1 >= Y0) & (by0 <= Y1);
            bal[j] = __ballot_sync(0xffffffffu, hit[j]);
            if (lane == 0) s_wcnt[par][j][warp] = __popc(bal[j]);
        }
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            const int i1 = top - kScanChunk - 1 - (j * kThreads + tid);
            nbox[j] = (i1 >= 0) ? __ldg(boxb + i1) : make_uint2(0xffff7fffu, 0xffff7fffu);
        }
        __syncthreads();
        int run = cnt;
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int c = s_wcnt[par][j][w];
                pre += (w < warp) ? c : 0;
                tot += c;
            }
            if (hit[j]) s_idx[run + pre + __popc(bal[j] & (lanebit - 1u))] = top - 1 - (j * kThreads + tid);
            run += tot;
        }
        cnt = run;
        par ^= 1;  // the next round's counts go to the other buffer: one barrier per round
        if (cnt > kListCap - kScanChunk || top <= kScanChunk) {
            __syncthreads();
            for (int e = tid; e < cnt; e += kThreads) {
                const int i = s_idx[e];
                const uint2 box = __ldg(boxb + i);
                const float4 *src = recb + (int64_t)i * 3;
                float4 q2 = __ldg(src + 2);
                q2.y = __uint_as_float(lane_mask((int)(short)(box.x & 0xffff), (int)box.x >> 16, X0));
                q2.z = __uint_as_float(row_code((int)(short)(box.y & 0xffff), (int)box.y >> 16, Y0,
                                                q2.w < 0.0f));
                cp_async16(&s_list[e * 3 + 0], src + 0);  // LDGSTS: no register staging
                cp_async16(&s_list[e * 3 + 1], src + 1);
                s_list[e * 3 + 2] = q2;
            }
            cp_async_wait_all();
            __syncthreads();
            if (live) live = composite_list<kStats>(s_list, cnt, lanebit, band_sel, Xf, Ybf, work);
            cnt = 0;
            // every band opaque: the rest of the genome is hidden behind what is drawn
            if (__syncthreads_and(!live)) break;
        }
    }
#else
    for (int top = N; top > 0; top -= kScanChunk) {
        bool hit[kScanPerThread];
        unsigned bal[kScanPerThread];
        int idx[kScanPerThread];
        int bx0[kScanPerThread], bx1[kScanPerThread], by0[kScanPerThread], by1[kScanPerThread];
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            idx[j] = top - 1 - (j * kThreads + tid);
#if GGS_BOX_PREFETCH
            const uint2 box = nbox[j];
            {
#else
            hit[j] = false;
            bx0[j] = bx1[j] = by0[j] = by1[j] = 0;
            if (idx[j] >= 0) {
                const uint2 box = __ldg(boxb + idx[j]);
#endif
                bx0[j] = (int)(short)(box.x & 0xffff);
                bx1[j] = (int)box.x >> 16;
                by0[j] = (int)(short)(box.y & 0xffff);
                by1[j] = (int)box.y >> 16;
                hit[j] = (bx1[j] >= X0) & (bx0[j] <= X1) & (by1[j] >= Y0) & (by0[j] <= Y1);
            }
            bal[j] = __ballot_sync(0xffffffffu, hit[j]);
            if (lane == 0) s_wcnt[j][warp] = __popc(bal[j]);
        }
#if GGS_BOX_PREFETCH
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            const int i1 = top - kScanChunk - 1 - (j * kThreads + tid);
            nbox[j] = (i1 >= 0) ? __ldg(boxb + i1) : make_uint2(0xffff7fffu, 0xffff7fffu);
        }
#endif
        __syncthreads();
        int run = cnt;
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            int pre = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int c = s_wcnt[j][w];
                pre += (w < warp) ? c : 0;
                tot += c;
            }
            if (hit[j]) {
                const int pos = run + pre + __popc(bal[j] & (lanebit - 1u));
                const float4 *src = recb + (int64_t)idx[j] * 3;
                float4 q2 = __ldg(src + 2);
                q2.y = __uint_as_float(lane_mask(bx0[j], bx1[j], X0));
                q2.z = __uint_as_float(row_code(by0[j], by1[j], Y0, q2.w < 0.0f));
                cp_async16(&s_list[pos * 3 + 0], src + 0);  // LDGSTS: no register staging
                cp_async16(&s_list[pos * 3 + 1], src + 1);
                s_list[pos * 3 + 2] = q2;
            }
            run += tot;
        }
        cnt = run;
        cp_async_wait_all();
        __syncthreads();
        if (cnt > kListCap - kScanChunk || top <= kScanChunk) {
            if (live) live = composite_list<kStats>(s_list, cnt, lanebit, band_sel, Xf, Ybf, work);
            cnt = 0;
            // every band opaque: the rest of the genome is hidden behind what is drawn
            if (__syncthreads_and(!live)) break;
        }
    }

#endif

    // Epilogue: add the background through the remaining transmittance (render.py:236-237),
    // clamp (render.py:252), optional image store, squared error (fitness.py:16-31).
    float num = 0.0f, den = 0.0f;
    const bool want_fit = (target != nullptr);
    float prr[kPairs][2], pgg[kPairs][2], pbb[kPairs][2], ptt[kPairs][2];
#define GGS_READ_ALL(k) GGS_PX_READ(k, prr[k], pgg[k], pbb[k], ptt[k])
    GGS_PAIRS(GGS_READ_ALL)
#undef GGS_READ_ALL
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
        const int Y = Yb + i;
        const float *pr = prr[i >> 1], *pg = pgg[i >> 1], *pb = pbb[i >> 1], *pt = ptt[i >> 1];
        if (X < W && Y < H) {
            const float tr = pt[i & 1];
            const float cr = clamp01(fmaf(tr, bg_r, pr[i & 1]));
            const float cg = clamp01(fmaf(tr, bg_g, pg[i & 1]));
            const float cb = clamp01(fmaf(tr, bg_b, pb[i & 1]));
            const int64_t p = (int64_t)Y * W + X;
            if (images != nullptr) {
                const int64_t at = ((int64_t)b * H * W + p) * 3;
                if (image_u8) {  // (img * 255).astype(uint8): truncation (utils.py:57)
                    unsigned char *o = reinterpret_cast<unsigned char *>(images) + at;
                    o[0] = (unsigned char)(cr * 255.0f);
                    o[1] = (unsigned char)(cg * 255.0f);
                    o[2] = (unsigned char)(cb * 255.0f);
                } else {
                    float *o = images + at;
                    o[0] = cr;
                    o[1] = cg;
                    o[2] = cb;
                }
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
    if (kStats && lane == 0) {
        // pixel-splat pairs actually evaluated: 64 lanes-rows per blended row pair
        atomicAdd(stats + 0, (unsigned long long)work[0]);
        atomicAdd(stats + 1, (unsigned long long)work[1]);
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
                          float *d_fitness, void *d_images, int image_u8,
                          unsigned long long *d_stats, cudaStream_t stream)
{
    if (B <= 0) return cudaSuccess;
    const int ntx = tiles_x(W), ntiles = ntx * tiles_y(H);
    const int64_t grid = (int64_t)B * ntiles;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (d_stats != nullptr)
        return launch_kernel(raster_kernel<true>, (unsigned)grid, kThreads, 0, stream, ws.rec, ws.aabb,
                             N, H, W, ntx, ntiles, bg[0], bg[1], bg[2], d_target, d_mask, mode, beta,
                             static_cast<float *>(d_images), image_u8, ws.partial, ws.counter,
                             d_fitness, d_stats);
    return launch_kernel(raster_kernel<false>, (unsigned)grid, kThreads, 0, stream, ws.rec, ws.aabb, N,
                         H, W, ntx, ntiles, bg[0], bg[1], bg[2], d_target, d_mask, mode, beta,
                         static_cast<float *>(d_images), image_u8, ws.partial, ws.counter, d_fitness,
                         nullptr);
}

}  // namespace ggs
