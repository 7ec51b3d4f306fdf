// One GA breeding step for the whole population in one launch ("next" #1):
// tournament selection, uniform row crossover, per-gene Gaussian mutation with the "at least
// one gene per group" rule, projection onto the legal genome box and the size-ordered splat
// swap -- the reference's
//   tournament_selection   modules/genetic.py:8-14     (+ shuffle/pairing, algorithm.py:87-100)
//   crossover_uniform      modules/genetic.py:17-21
//   mutate_individual      modules/genetic.py:32-92    (incl. _ensure_one_true :24-29)
//   clamp_genome           modules/utils.py:36-45
// which it runs per individual in Python (about 40 launches and 4-6 .item() syncs each).
// Here one CTA breeds one offspring pair; the population tensor never leaves the device.
//
// The job is to read two parents and write two children (36 B per splat each): HBM-bound if the
// arithmetic stays out of the way.  Three things keep it there:
//  * rows move through shared memory 256 at a time with element-wise, fully coalesced loads and
//    stores (a 9-float row per thread would touch every 128-byte line nine times) and are worked
//    on in place there (row stride 9 floats: conflict-free);
//  * randomness is spent only where it is used: one Philox call per (child, splat) yields the
//    seven Bernoulli flags (16-bit uniforms) and the crossover coin, and Gaussian noise is drawn
//    per (row, gene group) *item* -- the mutated groups of a chunk are compacted into a list and
//    processed densely, so a mutation rate of 10 % costs 10 % of the Box-Muller work instead
//    of a divergent branch that nearly every warp takes;
//  * the size-ordered swap scans the children while they are still in shared memory instead of
//    reading them back.
// Genomes with extra columns (cols > 9) or too many splats for the mask cache take a direct
// path (thread per row, global memory); both paths use the same counters and produce the same bits.
//
// Randomness is counter based (Philox4x32-10 keyed by the seed, counters = generation, child,
// splat, stream), so a step is reproducible for a given (seed, generation) and independent of
// the launch geometry.  The operators draw from the same distributions as the reference; the
// random streams necessarily differ (GA trajectory parity is not a goal).
//
// propose_kernel (below) is the same operator specialised for simulated annealing: children
// of ONE parent, no crossover, one CTA per child with the whole child in shared memory, and the
// decode of the finished rows (ggs_decode_math.cuh) in the same launch -- a sequential SA try is
// then propose -> raster -> Metropolis instead of breed -> decode -> raster -> Metropolis.  It
// uses the same counters and the same helpers, so it produces the same bits as breed_kernel.
// This file is compiled with -fmad=false because of the decode arithmetic.
#include <math.h>

#include "ggs_decode_math.cuh"

namespace ggs {
namespace {

constexpr int kBreedThreads = 256;
constexpr int kProposeThreads = 512;
constexpr int kProposeMaxSplats = 4096;  // the child lives in shared memory: 36 B + 1 B per splat
constexpr int kGroups = 5;  // xy, log-scales, theta, rgb, alpha
constexpr size_t kStageBytes = (size_t)2 * 2 * kBreedThreads * 9 * sizeof(float);

struct U4 {
    unsigned x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                            unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return {c0, c1, c2, c3};
}

// 4-byte asynchronous global -> shared copy (LDGSTS): the next chunk of parent rows streams in
// while the current one is being worked on.
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float u01(unsigned v) { return (float)(v >> 8) * (1.0f / 16777216.0f); }

// Two standard normals from two 32-bit words (Box-Muller).
__device__ __forceinline__ float2 normal2(unsigned a, unsigned b)
{
    // Hardware approximations (lg2 / sin / cos, absolute error ~2^-21): mutation noise needs the
    // right distribution, not last-bit accuracy, and the full-accuracy library calls were the
    // largest single item of the kernel's instruction count.
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.28318530717958647692f * (u01(b) - 0.5f), &s, &c);     // angle in [-pi, pi)
    return make_float2(r * c, r * s);
}

enum Stream : unsigned { kSelect = 1, kCross = 2, kMask = 3, kNoise = 4 /* +group */,
                         kForce = 16, kSwapIdx = 17, kSwapScore = 18 };

struct BreedParams {
    const float *pop;
    const float *fitness;
    float *off;
    int P, N, cols;
    int n_children;  // children to produce: the first n_children of the P the step defines
    int tour_k;
    float cxpb;
    unsigned mut_thr;  // mutpb as a 16-bit threshold: flag = (16-bit uniform < mut_thr)
    float s_xy, s_alog, s_blog, s_theta, s_rgb, s_alpha;
    float log_lo, log_hi;
    unsigned seed_lo, seed_hi, gen;
};

// One Philox call per (child, splat): seven Bernoulli(mutpb) flags from 16-bit uniforms
//   bit0,1 = x,y   bit2,3 = log sx, log sy   bit4 = theta   bit5 = rgb flag   bit6 = alpha flag
// (genetic.py:42-50) and, in bit 7, a fair coin: the row's crossover side (genetic.py:18), which
// is read from the pair's first child.
constexpr unsigned kFlagBits = 0x7fu, kCoinBit = 0x80u;
__device__ __forceinline__ unsigned gene_masks(const BreedParams &q, unsigned child, unsigned n)
{
    const U4 a = philox4x32_10(q.gen, child, n, kMask, q.seed_lo, q.seed_hi);
    unsigned m = 0;
    m |= ((a.x & 0xffffu) < q.mut_thr) ? 1u : 0u;
    m |= ((a.x >> 16) < q.mut_thr) ? 2u : 0u;
    m |= ((a.y & 0xffffu) < q.mut_thr) ? 4u : 0u;
    m |= ((a.y >> 16) < q.mut_thr) ? 8u : 0u;
    m |= ((a.z & 0xffffu) < q.mut_thr) ? 16u : 0u;
    m |= ((a.z >> 16) < q.mut_thr) ? 32u : 0u;
    m |= ((a.w & 0xffffu) < q.mut_thr) ? 64u : 0u;
    m |= ((a.w >> 16) & 1u) ? kCoinBit : 0u;
    return m;
}

// Which gene groups of a row mutate: bit g set iff any flag of group g is set.
__device__ __forceinline__ unsigned group_bits(unsigned m)
{
    return ((m & 3u) ? 1u : 0u) | ((m & 12u) ? 2u : 0u) | ((m & 16u) ? 4u : 0u) |
           ((m & 32u) ? 8u : 0u) | ((m & 64u) ? 16u : 0u);
}

// Gaussian mutation of gene group `grp` of one row (genetic.py:61-72); `row` is 9 floats in any
// address space.  The noise is a pure function of (generation, child, splat, group).
template <typename Row>
__device__ __forceinline__ void mutate_group(const BreedParams &q, unsigned child, unsigned n,
                                             int grp, unsigned m, Row &&row)
{
    const U4 r = philox4x32_10(q.gen, child, n, kNoise + (unsigned)grp, q.seed_lo, q.seed_hi);
    const float2 z = normal2(r.x, r.y);
    if (grp == 0) {
        if (m & 1u) row[0] += z.x * q.s_xy;
        if (m & 2u) row[1] += z.y * q.s_xy;
    } else if (grp == 1) {
        if (m & 4u) row[2] += z.x * q.s_alog;
        if (m & 8u) row[3] += z.y * q.s_blog;
    } else if (grp == 2) {
        row[4] += z.x * q.s_theta;
    } else if (grp == 3) {  // one flag for the three colour channels, independent noise
        const float2 z2 = normal2(r.z, r.w);
        row[5] += z.x * q.s_rgb;
        row[6] += z.y * q.s_rgb;
        row[7] += z2.x * q.s_rgb;
    } else {
        row[8] += z.x * q.s_alpha;
    }
}

__device__ __forceinline__ float wrap_angle(float t)
{
    // (t + pi) mod 2pi - pi (genetic.py:66): floor form; an angle already in range passes
    // through floor(...) = 0 exactly as it does through fmod.
    const float kPi = 3.14159265358979323846f, k2Pi = 6.28318530717958647692f;
    float v = t + kPi;
    v -= k2Pi * floorf(v * (1.0f / k2Pi));
    if (v < 0.0f) v += k2Pi;
    if (v >= k2Pi) v -= k2Pi;
    return v - kPi;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// Projection onto the legal box (utils.py:36-45).
__device__ __forceinline__ void project_row(const BreedParams &q, float (&g)[9])
{
    g[0] = clampf(g[0], 0.0f, 1.0f);
    g[1] = clampf(g[1], 0.0f, 1.0f);
    g[2] = clampf(g[2], q.log_lo, q.log_hi);
    g[3] = clampf(g[3], q.log_lo, q.log_hi);
    g[4] = wrap_angle(g[4]);
#pragma unroll
    for (int k = 5; k < 9; ++k) g[k] = clampf(g[k], 0.0f, 255.0f);
}

// Swap candidate j of a child (a later, bigger splat): uniform choice = arg max of iid scores.
// The score is a 32-bit integer hash (murmur3 finaliser) of j under a per-child Philox salt: a
// full Philox call per candidate cost more than everything else done to the row.
__device__ __forceinline__ unsigned long long swap_key(unsigned salt, int j)
{
    unsigned h = (unsigned)j * 0x9E3779B9u + salt;
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return ((unsigned long long)(h | 1u) << 32) | (unsigned)j;
}

// kStaged: cols == 9 and the mask cache fits; dynamic shared memory =
//   2 buffers x 2 parents x (kBreedThreads x 9) floats of row staging + 2 x N bytes of masks.
template <bool kStaged>
__global__ void __launch_bounds__(kBreedThreads) breed_kernel(BreedParams q)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    float *s_stage = reinterpret_cast<float *>(s_dyn);   // [buffer][parent / child][row][9]
    unsigned char *s_mask = s_dyn + kStageBytes;           // [2][N]
    __shared__ unsigned short s_items[2 * kBreedThreads * kGroups];  // child << 11 | row << 3 | group
    __shared__ unsigned char s_cmask[2][kBreedThreads];              // the chunk's masks, forced bits in
    __shared__ int s_nitems;
    __shared__ int s_parent[2];
    __shared__ int s_docx;
    __shared__ int s_count[2][4];   // mutated genes per child and group (xy, ab, theta, colour)
    __shared__ int s_force[2][4];   // forced element when a group came out empty, else -1
    __shared__ unsigned long long s_best[2];
    __shared__ int s_swap_i[2];
    __shared__ unsigned s_swap_salt[2];
    __shared__ float s_size_i[2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int pair = blockIdx.x;
    const int nchild = (2 * pair + 1 < q.n_children) ? 2 : 1;
    pdl_wait();
    pdl_trigger();

    if (tid == 0) {
        // two independent tournaments: a shuffled list of iid tournament winners paired up
        // (algorithm.py:87-96) is a sequence of iid pairs
        for (int side = 0; side < 2; ++side) {
            int best = -1;
            float bf = 0.0f;
            for (int d = 0; d < q.tour_k; ++d) {
                const U4 r = philox4x32_10(q.gen, pair, side * 1024 + d, kSelect, q.seed_lo, q.seed_hi);
                const int i = (int)__umulhi(r.x, (unsigned)q.P);
                const float f = q.fitness[i];
                if (best < 0 || f < bf) {
                    best = i;
                    bf = f;
                }
            }
            s_parent[side] = best;
        }
        const U4 r = philox4x32_10(q.gen, pair, 0, kCross, q.seed_lo, q.seed_hi);
        s_docx = (u01(r.x) < q.cxpb) ? 1 : 0;
    }
    if (tid < 8) {
        (&s_count[0][0])[tid] = 0;
        (&s_force[0][0])[tid] = -1;
    }
    if (tid < 2) {
        s_best[tid] = 0ull;
        // the splat that may be swapped back: uniform in [0, N-2] (genetic.py:76)
        const U4 r = philox4x32_10(q.gen, 2 * pair + tid, 0, kSwapIdx, q.seed_lo, q.seed_hi);
        s_swap_i[tid] = (q.N >= 2) ? (int)__umulhi(r.x, (unsigned)(q.N - 1)) : 0;
        s_swap_salt[tid] = r.y;
    }
    __syncthreads();

    const float *pa = q.pop + (int64_t)s_parent[0] * q.N * q.cols;
    const float *pb = q.pop + (int64_t)s_parent[1] * q.N * q.cols;
    // Stream chunk `n0` of both parents into staging buffer `buf`: element-wise, so a warp
    // moves 128 contiguous bytes per instruction whatever the row alignment.
    // 16-byte copies when every row block starts on a 16-byte boundary (N a multiple of 4 and
    // aligned tensors; a chunk is 256 rows = 9,216 B): a quarter of the copy instructions.
    const bool vec = ((q.N & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(q.pop) | reinterpret_cast<uintptr_t>(q.off)) & 15u) == 0;
    auto prefetch = [&](int n0, int buf) {
        const int nfl = min(kBreedThreads, q.N - n0) * 9;
        float *dst = s_stage + buf * (2 * kBreedThreads * 9);
        const float *ga = pa + (int64_t)n0 * 9, *gb = pb + (int64_t)n0 * 9;
        if (vec) {  // nfl is a multiple of 4 as well (rows is: N and the chunk size are)
            for (int i = 4 * tid; i < nfl; i += 4 * kBreedThreads) {
                cp_async16(dst + i, ga + i);
                cp_async16(dst + kBreedThreads * 9 + i, gb + i);
            }
        } else {
            for (int i = tid; i < nfl; i += kBreedThreads) {
                cp_async4(dst + i, ga + i);
                cp_async4(dst + kBreedThreads * 9 + i, gb + i);
            }
        }
    };
    if (kStaged) prefetch(0, 0);  // lands while the masks are drawn

    // ---- pass 1: draw the masks; how many genes of each group mutate ("at least one" rule)
    for (int c = 0; c < nchild; ++c) {
        const unsigned child = 2 * pair + c;
        int cnt[4] = {0, 0, 0, 0};
        for (int n = tid; n < q.N; n += kBreedThreads) {
            const unsigned m = gene_masks(q, child, n);
            if (kStaged) s_mask[c * q.N + n] = (unsigned char)m;
            cnt[0] += __popc(m & 3u);
            cnt[1] += __popc(m & 12u);
            cnt[2] += __popc(m & 16u);
            cnt[3] += __popc(m & 96u);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            int v = cnt[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(&s_count[c][g], v);
        }
    }
    __syncthreads();
    if (tid < 8) {
        // an empty group gets one uniformly chosen element forced on (genetic.py:24-29, 51-59)
        const int c = tid >> 2, g = tid & 3;
        if (c < nchild && s_count[c][g] == 0) {
            const int width = (g == 2) ? 1 : 2;
            const U4 r = philox4x32_10(q.gen, 2 * pair + c, g, kForce, q.seed_lo, q.seed_hi);
            s_force[c][g] = (int)__umulhi(r.x, (unsigned)(q.N * width));
        }
    }
    __syncthreads();

    // ---- pass 2: crossover, mutation, projection
    const int f[2][4] = {{s_force[0][0], s_force[0][1], s_force[0][2], s_force[0][3]},
                         {s_force[1][0], s_force[1][1], s_force[1][2], s_force[1][3]}};
    auto forced = [&](int c, int n, unsigned m) {
        m &= kFlagBits;
        if (f[c][0] >= 0 && (f[c][0] >> 1) == n) m |= 1u << (f[c][0] & 1);
        if (f[c][1] >= 0 && (f[c][1] >> 1) == n) m |= 4u << (f[c][1] & 1);
        if (f[c][2] >= 0 && f[c][2] == n) m |= 16u;
        if (f[c][3] >= 0 && (f[c][3] >> 1) == n) m |= 32u << (f[c][3] & 1);
        return m;
    };
    const int swap_i[2] = {s_swap_i[0], s_swap_i[1]};
    const unsigned swap_salt[2] = {s_swap_salt[0], s_swap_salt[1]};
    unsigned long long best[2] = {0ull, 0ull};  // this thread's best swap candidate per child

    if (kStaged) {
        float *off[2] = {q.off + (int64_t)(2 * pair) * q.N * 9, q.off + (int64_t)(2 * pair + 1) * q.N * 9};
        for (int n0 = 0, buf = 0; n0 < q.N; n0 += kBreedThreads, buf ^= 1) {
            const int rows = min(kBreedThreads, q.N - n0), nfl = rows * 9;
            float *s_rows[2] = {s_stage + buf * (2 * kBreedThreads * 9),
                                s_stage + buf * (2 * kBreedThreads * 9) + kBreedThreads * 9};
            if (tid == 0) s_nitems = 0;
            cp_async_wait_all();
            __syncthreads();  // this chunk has landed; the other buffer was stored and is free
            if (n0 + kBreedThreads < q.N) prefetch(n0 + kBreedThreads, buf ^ 1);

            // (a) crossover side of the row, masks, list of (row, group) items that need noise
            unsigned gm[2] = {0u, 0u};
            if (tid < rows) {
                const int n = n0 + tid;
                const unsigned m0 = s_mask[n];
                if (s_docx && !(m0 & kCoinBit)) {  // child 0 takes this row from parent b
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const float va = s_rows[0][tid * 9 + k];
                        s_rows[0][tid * 9 + k] = s_rows[1][tid * 9 + k];
                        s_rows[1][tid * 9 + k] = va;
                    }
                }
                for (int c = 0; c < nchild; ++c) {
                    const unsigned m = forced(c, n, c ? s_mask[q.N + n] : m0);
                    s_cmask[c][tid] = (unsigned char)m;
                    gm[c] = group_bits(m);
                }
            }
            const int mine = __popc(gm[0]) + __popc(gm[1]);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = 0;
            if (lane == 31 && incl) base = atomicAdd(&s_nitems, incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            int pos = base + incl - mine;
            for (int c = 0; c < 2; ++c)
                for (unsigned g = gm[c]; g; g &= g - 1u)
                    s_items[pos++] = (unsigned short)((c << 11) | (tid << 3) | (__ffs(g) - 1));
            __syncthreads();

            // (b) the noise, one item per thread
            for (int i = tid; i < s_nitems; i += kBreedThreads) {
                const unsigned it = s_items[i];
                const int c = it >> 11, t = (it >> 3) & 255, grp = it & 7;
                mutate_group(q, 2 * pair + c, n0 + t, grp, s_cmask[c][t], s_rows[c] + t * 9);
            }
            __syncthreads();

            // (c) projection, in place; the reference row of the swap announces its size
            if (tid < rows) {
                for (int c = 0; c < nchild; ++c) {
                    float g[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) g[k] = s_rows[c][tid * 9 + k];
                    project_row(q, g);
#pragma unroll
                    for (int k = 0; k < 9; ++k) s_rows[c][tid * 9 + k] = g[k];
                    if (n0 + tid == swap_i[c]) s_size_i[c] = g[2] + g[3];  // log(sigma_x * sigma_y)
                }
            }
            __syncthreads();

            // (d) later, bigger splats are swap candidates (genetic.py:77-83); store the chunk
            if (tid < rows && q.N >= 2) {
                for (int c = 0; c < nchild; ++c) {
                    const int j = n0 + tid;
                    if (j > swap_i[c] &&
                        s_rows[c][tid * 9 + 2] + s_rows[c][tid * 9 + 3] > s_size_i[c]) {
                        const unsigned long long key = swap_key(swap_salt[c], j);
                        best[c] = key > best[c] ? key : best[c];
                    }
                }
            }
            if (vec) {
                for (int i = 4 * tid; i < nfl; i += 4 * kBreedThreads) {
                    *reinterpret_cast<float4 *>(off[0] + (int64_t)n0 * 9 + i) =
                        *reinterpret_cast<const float4 *>(s_rows[0] + i);
                    if (nchild == 2)
                        *reinterpret_cast<float4 *>(off[1] + (int64_t)n0 * 9 + i) =
                            *reinterpret_cast<const float4 *>(s_rows[1] + i);
                }
            } else {
                for (int i = tid; i < nfl; i += kBreedThreads) {
                    off[0][(int64_t)n0 * 9 + i] = s_rows[0][i];
                    if (nchild == 2) off[1][(int64_t)n0 * 9 + i] = s_rows[1][i];
                }
            }
            __syncthreads();
        }
    } else {
        for (int n = tid; n < q.N; n += kBreedThreads) {
            const unsigned m0 = gene_masks(q, 2 * pair, n);
            const bool take_a = !s_docx || (m0 & kCoinBit);
            for (int c = 0; c < nchild; ++c) {
                const unsigned child = 2 * pair + c;
                const float *src = ((c == 0) == take_a ? pa : pb) + (int64_t)n * q.cols;
                float g[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) g[k] = __ldg(src + k);
                const unsigned m = forced(c, n, c ? gene_masks(q, child, n) : m0);
                const unsigned gmask = group_bits(m);
#pragma unroll
                for (int grp = 0; grp < kGroups; ++grp)
                    if (gmask & (1u << grp)) mutate_group(q, child, n, grp, m, g);
                project_row(q, g);
                float *dst = q.off + ((int64_t)child * q.N + n) * 9;
#pragma unroll
                for (int k = 0; k < 9; ++k) dst[k] = g[k];
            }
        }
        __syncthreads();  // the child rows written above are read back below by other threads
        // bring a bigger later splat forward (genetic.py:74-91): candidates from global memory
        if (q.N >= 2) {
            for (int c = 0; c < nchild; ++c) {
                const float *row = q.off + (int64_t)(2 * pair + c) * q.N * 9;
                const int i = swap_i[c];
                const float size_i = row[i * 9 + 2] + row[i * 9 + 3];
                for (int j = i + 1 + tid; j < q.N; j += kBreedThreads) {
                    if (row[j * 9 + 2] + row[j * 9 + 3] > size_i) {
                        const unsigned long long key = swap_key(swap_salt[c], j);
                        best[c] = key > best[c] ? key : best[c];
                    }
                }
            }
        }
    }

    // ---- pass 3: the swap itself
    if (q.N >= 2) {
        for (int c = 0; c < nchild; ++c) {
            unsigned long long b = best[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
                b = other > b ? other : b;
            }
            if (lane == 0 && b) atomicMax(&s_best[c], b);
        }
        __syncthreads();  // also orders the chunk stores above before the row exchange below
        if (tid < 9 * nchild) {
            const int c = tid / 9, k = tid - 9 * c;
            if (s_best[c] != 0ull) {
                float *row = q.off + (int64_t)(2 * pair + c) * q.N * 9;
                const int i = swap_i[c], j = (int)(s_best[c] & 0xffffffffull);
                const float a = row[i * 9 + k], b = row[j * 9 + k];
                row[i * 9 + k] = b;
                row[j * 9 + k] = a;
            }
        }
    }
}

// ---- simulated annealing: mutated copies of one parent, decoded in the same launch -----------
struct ProposeParams {
    BreedParams q;   // pop = the parent [N][cols] (P = 1), off = the children [n_children][N][9]
    float4 *rec;     // decoded records of the children: the workspace of their evaluation
    uint2 *aabb;
    int *counters;   // [n_children] raster tickets, cleared here (the decode launch would)
    int H, W;
    float k_sigma;
    ProposeJudge judge;  // the Metropolis step of the evaluation that preceded this proposal
};

__global__ void __launch_bounds__(kProposeThreads) propose_kernel(const __grid_constant__ ProposeParams p)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const BreedParams &q = p.q;
    float *s_rows = reinterpret_cast<float *>(s_dyn);                 // [N][9]
    unsigned char *s_mask = s_dyn + (size_t)q.N * 9 * sizeof(float);  // [N]
    __shared__ int s_count[4], s_force[4];
    __shared__ unsigned long long s_best;
    __shared__ int s_swap_i;
    __shared__ unsigned s_swap_salt;
    __shared__ float s_size_i;

    __shared__ int s_judged_cur, s_judged_best;

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned child = blockIdx.x;
    pdl_wait();
    pdl_trigger();

    // ---- judge the previous candidates (annealing.py:129-146); every CTA reaches the same verdict
    const ProposeJudge &J = p.judge;
    if (tid == 0) {
        int cur = -1, best = -1;
        if (J.tries > 0) {
            double e_cur = J.e_in[0], e_best = J.e_in[1];
            for (int k = 0; k < J.tries; ++k) {
                const double e_new = (double)J.energy[k];
                const double dE = e_new - e_cur;
                if (dE <= 0.0 || (J.temperature > 0.0 && J.uniform[k] < exp(-dE / J.temperature))) {
                    cur = k;
                    e_cur = e_new;
                }
                if (e_cur + 1e-12 < e_best) {
                    e_best = e_cur;
                    best = cur;
                }
            }
            if (blockIdx.x == 0) {
                J.e_out[0] = e_cur;
                J.e_out[1] = e_best;
                J.curve[0] = e_best;
                J.curve[1] = e_cur;
            }
        }
        s_judged_cur = cur;
        s_judged_best = best;
    }
    // (no barrier yet: the verdict is needed from pass 2 on, and the mask pass below does not
    // depend on it -- thread 0's loads overlap the other threads' Philox work)
    const int64_t row_floats = (int64_t)q.N * 9;
    if (q.n_children == 0) {  // a judging-only launch: move the rows and leave
        __syncthreads();
        if (J.tries > 0 && s_judged_best >= 0)
            for (int64_t i = tid; i < row_floats; i += kProposeThreads)
                J.best[i] = J.cand_prev[s_judged_best * row_floats + i];
        if (J.tries > 0 && s_judged_cur >= 0)
            for (int64_t i = tid; i < row_floats; i += kProposeThreads)
                J.current[i] = J.cand_prev[s_judged_cur * row_floats + i];
        return;
    }

    // the set-up belongs to warp 1, so that thread 0 (still judging) is not on its way
    if (tid >= 32 && tid < 36) {
        s_count[tid - 32] = 0;
        s_force[tid - 32] = -1;
    }
    if (tid == 36) {
        s_best = 0ull;
        const U4 r = philox4x32_10(q.gen, child, 0, kSwapIdx, q.seed_lo, q.seed_hi);
        s_swap_i = (q.N >= 2) ? (int)__umulhi(r.x, (unsigned)(q.N - 1)) : 0;
        s_swap_salt = r.y;
        p.counters[child] = 0;
    }

    // pass 1: the masks, and how many genes of each group mutate ("at least one" rule)
    int cnt[4] = {0, 0, 0, 0};
    for (int n = tid; n < q.N; n += kProposeThreads) {
        const unsigned m = gene_masks(q, child, n);
        s_mask[n] = (unsigned char)m;
        cnt[0] += __popc(m & 3u);
        cnt[1] += __popc(m & 12u);
        cnt[2] += __popc(m & 16u);
        cnt[3] += __popc(m & 96u);
    }
    __syncthreads();  // the counters are cleared, the verdict is in
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        int v = cnt[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&s_count[g], v);
    }
    __syncthreads();
    if (tid < 4 && s_count[tid] == 0) {  // an empty group gets one uniformly chosen element forced on
        const int width = (tid == 2) ? 1 : 2;
        const U4 r = philox4x32_10(q.gen, child, tid, kForce, q.seed_lo, q.seed_hi);
        s_force[tid] = (int)__umulhi(r.x, (unsigned)(q.N * width));
    }
    __syncthreads();

    // pass 2: mutation and projection of every row, into shared memory.  The parent is the
    // current state unless the verdict accepted a candidate of the previous batch; CTA 0 records
    // the accepted rows as the new current (and best) state on the way -- nobody reads `current`
    // when it is being replaced.
    const int f0 = s_force[0], f1 = s_force[1], f2 = s_force[2], f3 = s_force[3];
    const int swap_i = s_swap_i;
    const int j_cur = s_judged_cur, j_best = s_judged_best;
    const float *parent = q.pop;
    int parent_cols = q.cols;
    if (j_cur >= 0) {
        parent = J.cand_prev + j_cur * row_floats;
        parent_cols = 9;
    }
    const bool keep_cur = (blockIdx.x == 0 && j_cur >= 0);
    const bool keep_best = (blockIdx.x == 0 && j_best >= 0 && j_best == j_cur);
    if (blockIdx.x == 0 && j_best >= 0 && j_best != j_cur)  // the best was an earlier accepted try
        for (int64_t i = tid; i < row_floats; i += kProposeThreads)
            J.best[i] = J.cand_prev[j_best * row_floats + i];
    for (int n = tid; n < q.N; n += kProposeThreads) {
        const float *src = parent + (int64_t)n * parent_cols;
        float g[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = src[k];
        if (keep_cur) {
#pragma unroll
            for (int k = 0; k < 9; ++k) J.current[(int64_t)n * 9 + k] = g[k];
        }
        if (keep_best) {
#pragma unroll
            for (int k = 0; k < 9; ++k) J.best[(int64_t)n * 9 + k] = g[k];
        }
        unsigned m = s_mask[n] & kFlagBits;
        if (f0 >= 0 && (f0 >> 1) == n) m |= 1u << (f0 & 1);
        if (f1 >= 0 && (f1 >> 1) == n) m |= 4u << (f1 & 1);
        if (f2 >= 0 && f2 == n) m |= 16u;
        if (f3 >= 0 && (f3 >> 1) == n) m |= 32u << (f3 & 1);
        const unsigned gmask = group_bits(m);
#pragma unroll
        for (int grp = 0; grp < kGroups; ++grp)
            if (gmask & (1u << grp)) mutate_group(q, child, n, grp, m, g);
        project_row(q, g);
#pragma unroll
        for (int k = 0; k < 9; ++k) s_rows[n * 9 + k] = g[k];
        if (n == swap_i) s_size_i = g[2] + g[3];  // log(sigma_x * sigma_y)
    }
    __syncthreads();

    // bring a bigger later splat forward (genetic.py:74-91)
    if (q.N >= 2) {
        unsigned long long best = 0ull;
        const float size_i = s_size_i;
        const unsigned salt = s_swap_salt;
        for (int j = swap_i + 1 + tid; j < q.N; j += kProposeThreads) {
            if (s_rows[j * 9 + 2] + s_rows[j * 9 + 3] > size_i) {
                const unsigned long long key = swap_key(salt, j);
                best = key > best ? key : best;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0 && best) atomicMax(&s_best, best);
        __syncthreads();
        if (tid < 9 && s_best != 0ull) {
            const int j = (int)(s_best & 0xffffffffull);
            const float a = s_rows[swap_i * 9 + tid], b = s_rows[j * 9 + tid];
            s_rows[swap_i * 9 + tid] = b;
            s_rows[j * 9 + tid] = a;
        }
        __syncthreads();
    }

    // the child: genome rows for the caller, decoded records for the raster
    float *dst = q.off + (int64_t)child * q.N * 9;
    for (int i = tid; i < q.N * 9; i += kProposeThreads) dst[i] = s_rows[i];
    for (int n = tid; n < q.N; n += kProposeThreads) {
        float g[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = s_rows[n * 9 + k];
        const Decoded d = decode_chol(encode_axes(g), p.H, p.W, p.k_sigma);
        SplatRec r;
        uint2 box;
        make_record(d, r, box);
        const int64_t row = (int64_t)child * q.N + n;
        const float4 *rv = reinterpret_cast<const float4 *>(&r);
        p.rec[row * 3 + 0] = rv[0];
        p.rec[row * 3 + 1] = rv[1];
        p.rec[row * 3 + 2] = rv[2];
        p.aabb[row] = box;
    }
}

}  // namespace

bool propose_possible(int N, int cols) { return N >= 1 && N <= kProposeMaxSplats && cols >= 9; }

cudaError_t launch_propose(const float *d_parent, int N, int cols, int n_children, float *d_children,
                           float mutpb, const float sigma6[6], float log_lo, float log_hi,
                           uint64_t seed, uint32_t generation, const Workspace &ws, int H, int W,
                           float k_sigma, const ProposeJudge &judge, cudaStream_t stream)
{
    if (n_children <= 0 && judge.tries <= 0) return cudaSuccess;
    if (n_children < 0) n_children = 0;
    ProposeParams p;
    BreedParams &q = p.q;
    q.pop = d_parent;
    q.fitness = nullptr;
    q.off = d_children;
    q.P = 1;
    q.N = N;
    q.cols = cols;
    q.n_children = n_children;
    q.tour_k = 1;
    q.cxpb = 0.0f;
    const float thr = rintf(fminf(fmaxf(mutpb, 0.0f), 1.0f) * 65536.0f);
    q.mut_thr = (unsigned)thr;
    q.s_xy = sigma6[0];
    q.s_alog = sigma6[1];
    q.s_blog = sigma6[2];
    q.s_theta = sigma6[3];
    q.s_rgb = sigma6[4];
    q.s_alpha = sigma6[5];
    q.log_lo = log_lo;
    q.log_hi = log_hi;
    q.seed_lo = (unsigned)(seed & 0xffffffffu);
    q.seed_hi = (unsigned)(seed >> 32);
    q.gen = generation;
    p.rec = ws.rec;
    p.aabb = ws.aabb;
    p.counters = ws.counter;
    p.H = H;
    p.W = W;
    p.k_sigma = k_sigma;
    p.judge = judge;
    const size_t smem = (size_t)N * 9 * sizeof(float) + (size_t)N;
    static size_t granted[64] = {};
    int dev = 0;
    if (smem > (size_t)40 * 1024 && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 &&
        granted[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        granted[dev] = smem;
    }
    return launch_kernel(propose_kernel, n_children > 0 ? n_children : 1, kProposeThreads, smem, stream, p);
}

cudaError_t launch_breed(const float *d_pop, const float *d_fitness, int P, int N, int cols,
                         int n_children, float *d_offspring, int tour_k, float cxpb, float mutpb,
                         const float sigma6[6], float log_lo, float log_hi, uint64_t seed,
                         uint32_t generation, cudaStream_t stream)
{
    if (P <= 0 || N <= 0 || n_children <= 0) return cudaSuccess;
    BreedParams q;
    q.pop = d_pop;
    q.fitness = d_fitness;
    q.off = d_offspring;
    q.P = P;
    q.N = N;
    q.cols = cols;
    q.n_children = n_children;  // may exceed P: SA proposes many children of one individual
    q.tour_k = tour_k;
    q.cxpb = cxpb;
    const float thr = rintf(fminf(fmaxf(mutpb, 0.0f), 1.0f) * 65536.0f);
    q.mut_thr = (unsigned)thr;  // 0 .. 65536: Bernoulli(mutpb) to 1.5e-5
    q.s_xy = sigma6[0];
    q.s_alog = sigma6[1];
    q.s_blog = sigma6[2];
    q.s_theta = sigma6[3];
    q.s_rgb = sigma6[4];
    q.s_alpha = sigma6[5];
    q.log_lo = log_lo;
    q.log_hi = log_hi;
    q.seed_lo = (unsigned)(seed & 0xffffffffu);
    q.seed_hi = (unsigned)(seed >> 32);
    q.gen = generation;
    const size_t staged_bytes = kStageBytes + (size_t)2 * N;
    if (cols == 9 && staged_bytes <= (size_t)200 * 1024) {
        // opt in to more than the default dynamic shared memory only when it is needed, and only
        // when the limit already set on this device is too small (this is on the per-step path
        // of the GA / SA engines)
        static size_t granted[64] = {};
        int dev = 0;
        if (staged_bytes > (size_t)40 * 1024 && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 &&
            granted[dev] < staged_bytes) {
            cudaError_t e = cudaFuncSetAttribute(breed_kernel<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)staged_bytes);
            if (e != cudaSuccess) return e;
            granted[dev] = staged_bytes;
        }
        return launch_kernel(breed_kernel<true>, (q.n_children + 1) / 2, kBreedThreads, staged_bytes,
                             stream, q);
    }
    return launch_kernel(breed_kernel<false>, (q.n_children + 1) / 2, kBreedThreads, 0, stream, q);
}

}  // namespace ggs
