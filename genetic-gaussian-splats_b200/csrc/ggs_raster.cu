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
//
// Two kernels share the list building, the composite and the epilogue below:
//   raster_kernel        one CTA per (candidate, tile): the throughput path (populations);
//   raster_split_kernel  the latency path for very small batches (one SA try, a single frame, a
//                        handful of candidates), where one CTA per tile leaves most SMs idle: a
//                        thread-block CLUSTER of K = 2 / 4 / 8 CTAs shares a tile, CTA k composites
//                        the k-th segment of the genome front to back on its own, and the K partial
//                        (colour, transmittance) states -- "over" is associative:
//                        (C1,T1) o (C2,T2) = (C1 + T1 C2, T1 T2) -- are folded in genome order
//                        through distributed shared memory, each CTA finishing 1/K of the tile's
//                        pixels.  Its fused variant also decodes the genome rows itself (the
//                        arithmetic of ggs_decode_math.cuh): ONE launch per evaluation, but
//                        measured slower than decode + raster and therefore off by default.
// This file is compiled with -fmad=false (ggs_decode_math.cuh): every FMA below is explicit.
#include <cooperative_groups.h>

#include "ggs_decode_math.cuh"

namespace cg = cooperative_groups;

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
#ifndef GGS_OPT_NOFINALSYNC
#define GGS_OPT_NOFINALSYNC 1   // neutral at config 3, 4 % of a small launch
#endif
#ifndef GGS_OPT_EARLY_LOADS
#define GGS_OPT_EARLY_LOADS 0
#endif
#ifndef GGS_OPT_PREFETCH
#define GGS_OPT_PREFETCH 1
#endif
#ifndef GGS_EPI_BATCH
#define GGS_EPI_BATCH 4   // rows whose fitness inputs are loaded together in the epilogue (8 spills the pixel state)
#endif
#ifndef GGS_OPT_QY2
#define GGS_OPT_QY2 0
#endif
#ifndef GGS_SAT_EVERY
#define GGS_SAT_EVERY 8
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

#if defined(GGS_NAMED_REGS) && !GGS_NAMED_REGS   // only the C++ fallback of the pixel state uses these
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
#endif

// ---- pixel state ----------------------------------------------------------------------------
// Pair k of a thread holds rows (2k, 2k+1) of its column: accumulated colour (premultiplied,
// front to back) and transmittance.  GGS_PAIRS(M) expands M(0) ... M(kPairs-1).
//
// GGS_NAMED_REGS = 1 (default): the state lives in PTX registers declared ONCE per kernel
// (ggs_{r,g,b,t}<pair>, ggs_F, ggs_G) that only in-place PTX touches; as C++ values ptxas renamed
// the 64-bit accumulators out of place in most builds and paid ~20 MOVs per splat
// (tests/test_cpu_sass.py guards this).  The declaration has to stay in the kernel's entry block
// and every user has to be inlined into that kernel.
// GGS_NAMED_REGS = 0: the same operations on ordinary C++ values (struct PixelState), kept so
// that a toolchain that rejects or mis-scopes the hand-declared registers cannot brick the
// raster; tests/test_cpu_sass.py compiles it.
#ifndef GGS_NAMED_REGS
#define GGS_NAMED_REGS 1
#endif
#if GGS_ROWS == 8
#define GGS_PAIRS(M) M(0) M(1) M(2) M(3)
#else
#define GGS_PAIRS(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
#endif
#if GGS_NAMED_REGS
struct PixelState {};
#if GGS_ROWS == 8
#define GGS_PX_DECLARE()                                                                       \
    asm volatile(".reg .b64 ggs_r<4>, ggs_g<4>, ggs_b<4>, ggs_t<4>, ggs_F, ggs_G;");           \
    PixelState px
#else
#define GGS_PX_DECLARE()                                                                       \
    asm volatile(".reg .b64 ggs_r<8>, ggs_g<8>, ggs_b<8>, ggs_t<8>, ggs_F, ggs_G;");           \
    PixelState px
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
#else
struct PixelState {
    f2_t r[GGS_ROWS / 2], g[GGS_ROWS / 2], b[GGS_ROWS / 2], t[GGS_ROWS / 2], F, G;
};
#define GGS_PX_DECLARE() PixelState px
#define GGS_PX_INIT(k, ta_, tb_) \
    px.r[k] = 0ull;              \
    px.g[k] = 0ull;              \
    px.b[k] = 0ull;              \
    px.t[k] = pack2(ta_, tb_);
#define GGS_SET_F(f0_, f1_) px.F = pack2(f0_, f1_);
#define GGS_SET_G(g0_, g1_) px.G = pack2(g0_, g1_);
#define GGS_STEP_F() px.F = mul2(px.F, px.G);
#define GGS_STEP_G(H_) px.G = mul2(px.G, H_);
#define GGS_PX_BLEND(k)                         \
    {                                           \
        const f2_t w_ = mul2(px.F, px.t[k]);    \
        px.r[k] = fma2(w_, R2, px.r[k]);        \
        px.g[k] = fma2(w_, G2, px.g[k]);        \
        px.b[k] = fma2(w_, B2, px.b[k]);        \
        px.t[k] = sub2(px.t[k], w_);            \
    }
#define GGS_PX_READ_T(k, a_, b_) unpack2(px.t[k], a_, b_);
#define GGS_PX_READ(k, r_, g_, b_, t_)  \
    unpack2(px.r[k], r_[0], r_[1]);     \
    unpack2(px.g[k], g_[0], g_[1]);     \
    unpack2(px.b[k], b_[0], b_[1]);     \
    unpack2(px.t[k], t_[0], t_[1]);
#endif

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
__device__ __forceinline__ bool composite_list(PixelState &px, const float4 *__restrict__ list, int cnt,
                                               unsigned lanebit, unsigned band_sel, float Xf,
                                               float Ybf, unsigned (&work)[2])
{
    // the entry pointer advances by hand: with list[3 * s + k] ptxas of this build rebuilt the
    // address from the (uniform) counter with two FMA-pipe IMADs per entry
    const float4 *q = list;
    for (int s = 0; s < cnt; ++s, q += 3) {
        const float4 q2 = q[2];
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
#if GGS_OPT_EARLY_LOADS
        // issued with q2 so that their latency hides behind the band test; volatile asm, or the
        // compiler sinks the loads back below the branch
        float4 q0, q1;
        {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(q);
            asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(sa));
            asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];"
                         : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(sa));
        }
#endif
        const unsigned c = __byte_perm(__float_as_uint(q2.z), 0u, band_sel);  // this band's byte
        if (c == kBandMiss) continue;  // warp-uniform
#if !GGS_OPT_EARLY_LOADS
        const float4 q0 = q[0];
        const float4 q1 = q[1];
#endif
        const bool in_x = (__float_as_uint(q2.y) & lanebit) != 0u;
        const float qx = Xf - q0.x;
        const float t1 = q0.w * qx;                       // Bq*qx
        float t0 = fmaf(q0.z * qx, qx, q1.y);             // A*qx^2 + log2(alpha)
        t0 = in_x ? t0 : -INFINITY;                       // outside [x0,x1]: f = 2^-inf = 0
#if GGS_OPT_QY2
        const f2_t QY = add2(pack2(Ybf, Ybf + 1.0f), bcast2(-q0.y));  // one FADD2 (rows 0, 1 of the band)
#else
        const float dy = Ybf - q0.y;
        const f2_t QY = pack2(dy, dy + 1.0f);
#endif
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

// ------------------------------------------------------------------------------------------
// Kernel arguments (one struct for both kernels).
struct RasterArgs {
    // two-kernel path: records and cull boxes written by decode_kernel
    const float4 *rec;   // [B][N][3]
    const uint2 *aabb;   // [B][N]
    // fused path: the genomes themselves
    const float *genomes;  // [B][N][cols]
    int cols;
    float k_sigma;
    int N, H, W, ntx, ntiles;
    float bg_r, bg_g, bg_b;
    const float *target;  // [H][W][3] or NULL (render only)
    const float *mask;    // [H][W] or NULL
    int mode;
    float beta;
    float *images;  // [B][H][W][3] float or uint8, or NULL
    int image_u8;
    float2 *partial;  // [B][ntiles][split]
    int *counter;     // [B], zero at kernel start, zero again at kernel end
    float *fitness;   // [B]
    unsigned long long *stats;
    int split;  // CTAs per (candidate, tile): the cluster size of raster_split_kernel
    int prefetch_inputs;  // latency regime: pull the tile's target / mask lines into L1 up front
    int B;                // candidates of this launch
    PeerStores peers;  // fitness stores into the other GPUs' gathered vectors (ggs_peers.cu)
};

struct TileGeom {
    int b, t, X0, Y0, X1, Y1, X, Yb;
    float Xf, Ybf;
    unsigned lanebit, band_sel;
};

// CTA index -> (candidate, tile).  Candidate-major by default: the tiles of a candidate run
// together and share its records in L1 / L2.  `interior_first` (grids from two CTAs per SM to a few
// waves, where the LAST CTAs an SM is dealt decide the launch time): tile-major, the image's tiles
// from the centre outwards, so that the CTAs that start last -- or that land on the SMs holding
// one CTA more than the others -- are the cheap ones: a border tile lists about 60 % of the
// splats of an interior one, a corner tile a third.  Results do not depend on the order (partials
// are stored and summed by tile index).
template <bool kInteriorFirst = false>
__device__ __forceinline__ TileGeom tile_geometry(int cand_tile, int ntx, int ntiles, int lane, int warp,
                                                  int B = 1)
{
    TileGeom g;
    int tx, ty;
    if (kInteriorFirst) {
        const int nty = ntiles / ntx;
        const int r = cand_tile / B;
        g.b = cand_tile - r * B;
        centre_out_tile(r, ntx, nty, tx, ty);
    } else {
        g.b = cand_tile / ntiles;
        const int t = cand_tile - g.b * ntiles;
        ty = t / ntx;
        tx = t - ty * ntx;
    }
    g.t = ty * ntx + tx;
    g.X0 = tx * kTileW;
    g.Y0 = ty * kTileH;
    g.X1 = g.X0 + kTileW - 1;
    g.Y1 = g.Y0 + kTileH - 1;
    g.X = g.X0 + lane;
    g.Yb = g.Y0 + warp * kRowsPerThread;
    // Loop constants of the composite, pinned in registers by a (no-op) shuffle.
    g.Xf = __shfl_sync(0xffffffffu, (float)g.X, lane);
    g.Ybf = __shfl_sync(0xffffffffu, (float)g.Yb, lane);
    g.lanebit = __shfl_sync(0xffffffffu, 1u << lane, lane);
    g.band_sel = __shfl_sync(0xffffffffu, 0x4440u + (unsigned)warp, lane);  // PRMT: byte `warp`
    return g;
}

// Shared memory of a CTA (static, 26 KB): the staged list, the index list of the scan, the
// per-warp hit counts of two rounds, the fitness reduction scratch.  The arrays are declared one
// by one in the kernels (GGS_SMEM_DECLARE) and handed to the helpers as pointers: as members of
// one struct ptxas stopped keeping the composite loop's counter and list address in uniform
// registers (17 uniform-datapath operations per list entry became vector ones: 7 % slower).
struct RasterSmem {
    float4 *list;                          // [kListCap * 3]
    int *idx;                              // [kListCap]
    int (*wcnt)[kScanPerThread][kWarps];   // [2]
    float *red;                            // [2 * kWarps]
    int *last_;
};
#define GGS_SMEM_DECLARE()                                   \
    __shared__ float4 s_list[kListCap * 3];                  \
    __shared__ int s_wcnt[2][kScanPerThread][kWarps];        \
    __shared__ int s_idx[kListCap];                          \
    __shared__ float s_red[2 * kWarps];                      \
    __shared__ int s_last;                                   \
    const RasterSmem sm = {s_list, s_idx, s_wcnt, s_red, &s_last}

// Walk records [0, n) of `recb` / `boxb` from the last to the first (front to back) and blend
// the ones that touch the tile.  Slot j of a round maps thread `tid` to record
// top - 1 - (j*kThreads + tid): ascending (j, tid) is descending genome order, so the ordinary
// ballot compaction yields the order we need.  A round only compacts the INDICES of the hits
// (ordered, by ballot and a cross-warp prefix); the records are staged when the list is about
// to be composited, one list entry per thread, so the staging code runs ceil(cnt / kThreads)
// times per flush instead of once per (round, slot) with a few lanes active.
template <bool kStats>
__device__ __forceinline__ void scan_and_composite(PixelState &px, const float4 *__restrict__ recb,
                                                   const uint2 *__restrict__ boxb, int n,
                                                   const TileGeom &g, const RasterSmem &sm, int tid, int lane,
                                                   int warp, unsigned (&work)[2])
{
    bool live = true;  // warp-uniform: this band still has a non-opaque pixel
    int cnt = 0;
    uint2 nbox[kScanPerThread];  // the round's boxes, loaded one round ahead
#pragma unroll
    for (int j = 0; j < kScanPerThread; ++j) {
        const int i0 = n - 1 - (j * kThreads + tid);
        nbox[j] = (i0 >= 0) ? __ldg(boxb + i0) : make_uint2(0xffff7fffu, 0xffff7fffu);
    }
    int par = 0;
    for (int top = n; top > 0; top -= kScanChunk) {
        bool hit[kScanPerThread];
        unsigned bal[kScanPerThread];
#pragma unroll
        for (int j = 0; j < kScanPerThread; ++j) {
            const uint2 box = nbox[j];
            const int bx0 = (int)(short)(box.x & 0xffff), bx1 = (int)box.x >> 16;
            const int by0 = (int)(short)(box.y & 0xffff), by1 = (int)box.y >> 16;
            hit[j] = (bx1 >= g.X0) & (bx0 <= g.X1) & (by1 >= g.Y0) & (by0 <= g.Y1);
            bal[j] = __ballot_sync(0xffffffffu, hit[j]);
            if (lane == 0) sm.wcnt[par][j][warp] = __popc(bal[j]);
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
                const int c = sm.wcnt[par][j][w];
                pre += (w < warp) ? c : 0;
                tot += c;
            }
            if (hit[j]) sm.idx[run + pre + __popc(bal[j] & (g.lanebit - 1u))] = top - 1 - (j * kThreads + tid);
            run += tot;
        }
        cnt = run;
        par ^= 1;  // the next round's counts go to the other buffer: one barrier per round
        if (cnt > kListCap - kScanChunk || top <= kScanChunk) {
            __syncthreads();
            for (int e = tid; e < cnt; e += kThreads) {
                const int i = sm.idx[e];
                const uint2 box = __ldg(boxb + i);
                const float4 *src = recb + (int64_t)i * 3;
                float4 q2 = __ldg(src + 2);
                q2.y = __uint_as_float(lane_mask((int)(short)(box.x & 0xffff), (int)box.x >> 16, g.X0));
                q2.z = __uint_as_float(row_code((int)(short)(box.y & 0xffff), (int)box.y >> 16, g.Y0,
                                                q2.w < 0.0f));
                cp_async16(&sm.list[e * 3 + 0], src + 0);  // LDGSTS: no register staging
                cp_async16(&sm.list[e * 3 + 1], src + 1);
                sm.list[e * 3 + 2] = q2;
            }
            cp_async_wait_all();
            __syncthreads();
            if (live)
                live = composite_list<kStats>(px, sm.list, cnt, g.lanebit, g.band_sel, g.Xf, g.Ybf, work);
            cnt = 0;
#if GGS_OPT_NOFINALSYNC
            if (top <= kScanChunk) break;  // that was the last flush: no reason to wait for the other bands
#endif
            // every band opaque: the rest of the genome is hidden behind what is drawn
            if (__syncthreads_and(!live)) break;
        }
    }
}

// Fused decode (latency path): the CTA decodes rows [i_lo, i_lo + n) of the candidate's genome
// itself, kThreads per round in descending order, and stages the ones that touch the tile
// straight into the list (n <= kListCap, so there is a single composite).  Same arithmetic as
// decode_kernel (ggs_decode_math.cuh), hence the same records and the same image bits as the
// two-kernel path.
template <bool kAxes>
__device__ __forceinline__ int build_list_fused(const float *__restrict__ gb, int cols, int i_lo, int n,
                                                int H, int W, float k_sigma, const TileGeom &g,
                                                const RasterSmem &sm, int tid, int lane, int warp)
{
    int cnt = 0, par = 0;
    for (int top = n; top > 0; top -= kThreads) {
        const int rel = top - 1 - tid;
        bool hit = false;
        SplatRec r;
        int bx0 = 0, bx1 = 0, by0 = 0, by1 = 0;
        if (rel >= 0) {
            const float *src = gb + (int64_t)(i_lo + rel) * cols;
            float v[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = __ldg(src + c);
            const Chol ch = kAxes ? encode_axes(v) : load_chol(v);
            const Decoded d = decode_chol(ch, H, W, k_sigma);
            uint2 box;
            make_record(d, r, box);
            bx0 = (int)(short)(box.x & 0xffff), bx1 = (int)box.x >> 16;
            by0 = (int)(short)(box.y & 0xffff), by1 = (int)box.y >> 16;
            hit = (bx1 >= g.X0) & (bx0 <= g.X1) & (by1 >= g.Y0) & (by0 <= g.Y1);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) sm.wcnt[par][0][warp] = __popc(bal);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int c = sm.wcnt[par][0][w];
            pre += (w < warp) ? c : 0;
            tot += c;
        }
        if (hit) {
            const int e = cnt + pre + __popc(bal & (g.lanebit - 1u));
            sm.list[e * 3 + 0] = make_float4(r.cx, r.cy, r.A, r.Bq);
            sm.list[e * 3 + 1] = make_float4(r.Cq, r.la, r.r, r.g);
            sm.list[e * 3 + 2] = make_float4(r.b, __uint_as_float(lane_mask(bx0, bx1, g.X0)),
                                             __uint_as_float(row_code(by0, by1, g.Y0, r.h < 0.0f)), r.h);
        }
        cnt += tot;
        par ^= 1;
    }
    __syncthreads();
    return cnt;
}

// The fitness inputs of one pixel: target colour and weight (fitness.py:16-31).  Issued for all of
// a thread's pixels BEFORE any of them is used, so the epilogue pays one L2 round trip, not one
// per row (measured: a third of a small launch's time went into eight serial round trips).
struct PixelInputs {
    float tr, tg, tb, w;
};

__device__ __forceinline__ PixelInputs load_inputs(const RasterArgs &a, int X, int Y)
{
    PixelInputs in = {0.0f, 0.0f, 0.0f, 1.0f};
    if (a.target != nullptr && X < a.W && Y < a.H) {
        const int64_t p = (int64_t)Y * a.W + X;
        in.tr = __ldg(a.target + 3 * p + 0);
        in.tg = __ldg(a.target + 3 * p + 1);
        in.tb = __ldg(a.target + 3 * p + 2);
        if (a.mode != GGS_MODE_PLAIN) in.w = __ldg(a.mask + p);
    }
    return in;
}

// One finished pixel: background through the remaining transmittance (render.py:236-237), clamp
// (render.py:252), optional image store, squared error (fitness.py:16-31).
__device__ __forceinline__ void emit_pixel(const RasterArgs &a, int b, int X, int Y, float pr, float pg,
                                           float pb, float pt, const PixelInputs &in, float &num,
                                           float &den)
{
    const float cr = clamp01(fmaf(pt, a.bg_r, pr));
    const float cg = clamp01(fmaf(pt, a.bg_g, pg));
    const float cb = clamp01(fmaf(pt, a.bg_b, pb));
    if (a.images != nullptr) {
        const int64_t at = (((int64_t)b * a.H + Y) * a.W + X) * 3;
        if (a.image_u8) {  // (img * 255).astype(uint8): truncation (utils.py:57)
            unsigned char *o = reinterpret_cast<unsigned char *>(a.images) + at;
            o[0] = (unsigned char)(cr * 255.0f);
            o[1] = (unsigned char)(cg * 255.0f);
            o[2] = (unsigned char)(cb * 255.0f);
        } else {
            float *o = a.images + at;
            o[0] = cr;
            o[1] = cg;
            o[2] = cb;
        }
    }
    if (a.target != nullptr) {
        const float dr = cr - in.tr, dg = cg - in.tg, db = cb - in.tb;
        const float w = (a.mode == GGS_MODE_BOOST) ? fmaf(a.beta, clamp01(in.w), 1.0f) : in.w;
        num += (dr * dr) * w + (dg * dg) * w + (db * db) * w;
        den += w;
    }
}

// Block reduction of (num, den) to this CTA's partial, then the last CTA of the candidate
// (atomic ticket over its `per_cand` CTAs) combines all partials in index order, in double, and
// applies the mode formula: fitness is bit-reproducible.  Deliberately NOT inlined: it runs once
// per CTA, and with its body (and the peer stores) inside the kernel ptxas moved the composite
// loop's list address out of the uniform registers.
__device__ __noinline__ void reduce_and_finish(const RasterArgs &a, int b, int slot, int per_cand, float num,
                                               float den, const RasterSmem &sm, int tid, int lane, int warp)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (lane == 0) {
        sm.red[warp] = num;
        sm.red[kWarps + warp] = den;
    }
    __syncthreads();
    if (warp != 0) return;  // one warp finishes the CTA; the others free their slots now
    int ticket = 0;
    if (lane == 0) {
        float n = 0.0f, d = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            n += sm.red[w];
            d += sm.red[kWarps + w];
        }
        a.partial[(int64_t)b * per_cand + slot] = make_float2(n, d);  // by tile, whatever the CTA order
        // release our partial, acquire everybody else's: one acq_rel RMW instead of fence + atomic
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(ticket) : "l"(a.counter + b) : "memory");
    }
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket != per_cand - 1) return;
    // last CTA of this candidate: combine the partials in index order
    const float2 *pb = a.partial + (int64_t)b * per_cand;
    double n = 0.0, d = 0.0;
    for (int k = lane; k < per_cand; k += 32) {
        const float2 v = __ldcg(pb + k);
        n += (double)v.x;
        d += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if (lane == 0) {
        const double P = (double)a.H * (double)a.W;
        float fit;
        if (a.mode == GGS_MODE_PLAIN)
            fit = (float)(n / (3.0 * P));                            // fitness.py:19
        else if (a.mode == GGS_MODE_MASK)
            fit = (float)n / ((float)d + 1e-12f);                    // fitness.py:29-31
        else
            fit = (float)(n / (3.0 * P)) / ((float)(d / P) + 1e-12f);  // fitness.py:23-27
        a.fitness[b] = fit;
        a.counter[b] = 0;  // ready for the next launch on this workspace
        peer_publish(a.peers, b, fit, (int)(gridDim.x / (unsigned)per_cand));
    }
}

// The tile's slice of the target and the mask is what the epilogue waits for (one L2 round trip
// per row when it is read cold: half of a small launch's time).  It does not depend on the
// decode launch ahead of us, so every thread asks for one 128-byte line of it -- 32 rows x
// (3 lines of target + 1 of mask) -- BEFORE the grid dependency wait; by the epilogue it sits in L1.
__device__ __forceinline__ void prefetch_tile_inputs(const RasterArgs &a, const TileGeom &g, int tid)
{
#if !GGS_OPT_PREFETCH
    return;
#endif
    // only for grids of at most one wave: at config 3 the extra L2 -> L1 fills (16 KB per CTA next
    // to ~19 KB of boxes and records) cost 0.45 % and the epilogue's latency is hidden anyway
    if (!a.prefetch_inputs || a.target == nullptr || g.X0 >= a.W) return;
    const int px = min(kTileW, a.W - g.X0);  // pixels of a row inside the image
    const int part = tid & 3;
    for (int r = tid >> 2; r < kTileH; r += kThreads / 4) {
        const int row = g.Y0 + r;
        if (row >= a.H) break;
        const int64_t p = (int64_t)row * a.W + g.X0;
        const float *addr = nullptr;
        if (part < 3) {
            if (part * 32 < px * 3) addr = a.target + 3 * p + part * 32;
        } else if (a.mode != GGS_MODE_PLAIN) {
            addr = a.mask + p;
        }
        if (addr != nullptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(addr));
    }
}

// Transmittance starts at 1 inside the image and at 0 outside it: pixels beyond the image edge
// then take no colour and never keep a band from saturating.
#define GGS_T_INIT(k)                                                        \
    GGS_PX_INIT(k, (g.X < a.W && g.Yb + 2 * k < a.H) ? 1.0f : 0.0f,          \
                (g.X < a.W && g.Yb + 2 * k + 1 < a.H) ? 1.0f : 0.0f)

// ---- throughput path: one CTA per (candidate, tile) -------------------------------------------
// kInteriorFirst is a template parameter, not an argument: with both orders in one instance the
// throughput kernel (config 3) lost 1.7 % to the extra prologue code and its spills.
template <bool kStats, bool kInteriorFirst>
__global__ void __launch_bounds__(kThreads, GGS_MIN_BLOCKS) raster_kernel(const __grid_constant__ RasterArgs a)
{
    unsigned work[2] = {0u, 0u};  // kStats: row pairs blended on the recurrence / exact path
    GGS_SMEM_DECLARE();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileGeom g = tile_geometry<kInteriorFirst>(blockIdx.x, a.ntx, a.ntiles, lane, warp, a.B);

    GGS_PX_DECLARE();
    GGS_PAIRS(GGS_T_INIT)

    prefetch_tile_inputs(a, g, tid);  // inputs nobody ahead of us writes
    pdl_wait();  // the decode launch ahead of us has completed; nothing above reads its output
    pdl_trigger();
    scan_and_composite<kStats>(px, a.rec + (int64_t)g.b * a.N * 3, a.aabb + (int64_t)g.b * a.N, a.N, g, sm,
                               tid, lane, warp, work);

    float num = 0.0f, den = 0.0f;
    float prr[kPairs][2], pgg[kPairs][2], pbb[kPairs][2], ptt[kPairs][2];
#define GGS_READ_ALL(k) GGS_PX_READ(k, prr[k], pgg[k], pbb[k], ptt[k])
#if GGS_EPI_BATCH >= GGS_ROWS
    PixelInputs in[kRowsPerThread];
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) in[i] = load_inputs(a, g.X, g.Yb + i);
    GGS_PAIRS(GGS_READ_ALL)
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
        const int Y = g.Yb + i;
        if (g.X < a.W && Y < a.H)
            emit_pixel(a, g.b, g.X, Y, prr[i >> 1][i & 1], pgg[i >> 1][i & 1], pbb[i >> 1][i & 1],
                       ptt[i >> 1][i & 1], in[i], num, den);
    }
#else
    GGS_PAIRS(GGS_READ_ALL)
#pragma unroll
    for (int i0 = 0; i0 < kRowsPerThread; i0 += GGS_EPI_BATCH) {
        PixelInputs in[GGS_EPI_BATCH];
#pragma unroll
        for (int i = 0; i < GGS_EPI_BATCH; ++i) in[i] = load_inputs(a, g.X, g.Yb + i0 + i);
#pragma unroll
        for (int j = 0; j < GGS_EPI_BATCH; ++j) {
            const int i = i0 + j, Y = g.Yb + i;
            if (g.X < a.W && Y < a.H)
                emit_pixel(a, g.b, g.X, Y, prr[i >> 1][i & 1], pgg[i >> 1][i & 1], pbb[i >> 1][i & 1],
                           ptt[i >> 1][i & 1], in[j], num, den);
        }
    }
#endif
    if (kStats && lane == 0) {
        // pixel-splat pairs actually evaluated: 64 lanes-rows per blended row pair
        atomicAdd(a.stats + 0, (unsigned long long)work[0]);
        atomicAdd(a.stats + 1, (unsigned long long)work[1]);
    }
    if (a.target == nullptr) return;
    reduce_and_finish(a, g.b, g.t, a.ntiles, num, den, sm, tid, lane, warp);
}

// ---- latency path: a cluster of `split` CTAs per (candidate, tile) ----------------------------
// kDecode: 0 = records from decode_kernel, 1 = fused decode of axes-angle genomes, 2 = fused
// decode of Cholesky genomes.  blockIdx.x = (candidate * ntiles + tile) * split + k; CTA k owns
// genome rows [k*S, min(N, (k+1)*S)), S = ceil(N / split); the highest k is the front-most.
template <int kDecode, bool kInteriorFirst>
__global__ void __launch_bounds__(kThreads, GGS_MIN_BLOCKS) raster_split_kernel(const __grid_constant__ RasterArgs a)
{
    unsigned work[2] = {0u, 0u};
    GGS_SMEM_DECLARE();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.split;
    const int k = (int)(blockIdx.x % (unsigned)K);
    const TileGeom g = tile_geometry<kInteriorFirst>((int)(blockIdx.x / (unsigned)K), a.ntx, a.ntiles, lane, warp, a.B);
    const int S = (a.N + K - 1) / K;
    const int i_lo = min(k * S, a.N), n = min(a.N, i_lo + S) - i_lo;

    GGS_PX_DECLARE();
    GGS_PAIRS(GGS_T_INIT)

    prefetch_tile_inputs(a, g, tid);
    pdl_wait();
    pdl_trigger();
    if (kDecode == 0) {
        scan_and_composite<false>(px, a.rec + ((int64_t)g.b * a.N + i_lo) * 3,
                                  a.aabb + (int64_t)g.b * a.N + i_lo, n, g, sm, tid, lane, warp, work);
    } else {
        const int cnt = build_list_fused<kDecode == 1>(a.genomes + (int64_t)g.b * a.N * a.cols, a.cols,
                                                       i_lo, n, a.H, a.W, a.k_sigma, g, sm, tid, lane, warp);
        composite_list<false>(px, sm.list, cnt, g.lanebit, g.band_sel, g.Xf, g.Ybf, work);
    }

    // Publish this segment's state: exch[row][col] = (r, g, b, t), reusing the list's memory.
    static_assert(kTileH * kTileW <= kListCap * 3, "the tile's pixel states must fit the list buffer");
    __syncthreads();  // every warp is done reading the list
    float4 *exch = sm.list;
    {
        float prr[kPairs][2], pgg[kPairs][2], pbb[kPairs][2], ptt[kPairs][2];
        GGS_PAIRS(GGS_READ_ALL)
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i)
            exch[(warp * kRowsPerThread + i) * kTileW + lane] =
                make_float4(prr[i >> 1][i & 1], pgg[i >> 1][i & 1], pbb[i >> 1][i & 1], ptt[i >> 1][i & 1]);
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();

    // Fold the K states front to back for this CTA's share of the tile's rows, then finish
    // those pixels.  Segment K-1 holds the last splats of the genome: it is the front-most.
    float num = 0.0f, den = 0.0f;
    const int rows_per = kTileH / K;
    for (int p = tid; p < rows_per * kTileW; p += kThreads) {
        const int row = k * rows_per + p / kTileW, col = p % kTileW;
        const int X = g.X0 + col, Y = g.Y0 + row;
        // every load of this pixel -- the K remote states and the fitness inputs -- is issued
        // before the first use: one trip over the cluster network and one to L2, not K + 1
        float4 v[kMaxSplit];
#pragma unroll
        for (int s = 0; s < kMaxSplit; ++s)
            v[s] = (s < K) ? cluster.map_shared_rank(exch, s)[row * kTileW + col]
                           : make_float4(0.0f, 0.0f, 0.0f, 1.0f);  // the identity of the fold
        const PixelInputs in = load_inputs(a, X, Y);
        float cr = 0.0f, cgr = 0.0f, cb = 0.0f, t = 1.0f;
#pragma unroll
        for (int s = kMaxSplit - 1; s >= 0; --s) {
            if (s < K) {
                cr = fmaf(t, v[s].x, cr);
                cgr = fmaf(t, v[s].y, cgr);
                cb = fmaf(t, v[s].z, cb);
                t *= v[s].w;
            }
        }
        if (X < a.W && Y < a.H) emit_pixel(a, g.b, X, Y, cr, cgr, cb, t, in, num, den);
    }
    cluster.sync();  // nobody leaves while a neighbour still reads its shared memory
    if (a.target == nullptr) return;
    reduce_and_finish(a, g.b, g.t * K + k, a.ntiles * K, num, den, sm, tid, lane, warp);
}
#undef GGS_READ_ALL
#undef GGS_T_INIT

}  // namespace

int max_split_for(int N) { return N <= 0 ? 1 : kMaxSplit; }

bool fused_decode_possible(int N, int split)
{
    return split >= 1 && (N + split - 1) / split <= kListCap;
}

cudaError_t launch_raster(const RasterLaunch &q, cudaStream_t stream)
{
    if (q.B <= 0) return cudaSuccess;
    const int ntx = tiles_x(q.W), ntiles = ntx * tiles_y(q.H);
    const int split = q.split < 1 ? 1 : q.split;
    const int64_t grid = (int64_t)q.B * ntiles * split;
    if (grid > 0x7fffffffLL || kTileH % split != 0 || split > kMaxSplit) return cudaErrorInvalidConfiguration;
    RasterArgs a;
    a.rec = q.ws.rec;
    a.aabb = q.ws.aabb;
    a.genomes = q.d_genomes;
    a.cols = q.cols;
    a.k_sigma = q.k_sigma;
    a.N = q.N;
    a.H = q.H;
    a.W = q.W;
    a.ntx = ntx;
    a.ntiles = ntiles;
    a.bg_r = q.bg[0];
    a.bg_g = q.bg[1];
    a.bg_b = q.bg[2];
    a.target = q.d_target;
    a.mask = q.d_mask;
    a.mode = q.mode;
    a.beta = q.beta;
    a.images = static_cast<float *>(q.d_images);
    a.image_u8 = q.image_u8;
    a.partial = q.ws.partial;
    a.counter = q.ws.counter;
    a.fitness = q.d_fitness;
    a.stats = q.d_stats;
    a.split = split;
    a.prefetch_inputs = q.small_grid ? 1 : 0;
    a.B = q.B;
    const bool interior_first = q.interior_first && ntx >= 3 && ntiles / ntx >= 3;
    a.peers = q.peers;
    if (q.fused) {
        if (!fused_decode_possible(q.N, split)) return cudaErrorInvalidConfiguration;
        if (q.layout == GGS_LAYOUT_AXES_ANGLE)
            return launch_kernel_cluster(raster_split_kernel<1, false>, (unsigned)grid, kThreads, 0, split, stream, a);
        return launch_kernel_cluster(raster_split_kernel<2, false>, (unsigned)grid, kThreads, 0, split, stream, a);
    }
    if (split > 1) {
        if (interior_first)
            return launch_kernel_cluster(raster_split_kernel<0, true>, (unsigned)grid, kThreads, 0, split, stream, a);
        return launch_kernel_cluster(raster_split_kernel<0, false>, (unsigned)grid, kThreads, 0, split, stream, a);
    }
    if (q.d_stats != nullptr)
        return launch_kernel(raster_kernel<true, false>, (unsigned)grid, kThreads, 0, stream, a);
    if (interior_first)
        return launch_kernel(raster_kernel<false, true>, (unsigned)grid, kThreads, 0, stream, a);
    return launch_kernel(raster_kernel<false, false>, (unsigned)grid, kThreads, 0, stream, a);
}

}  // namespace ggs
