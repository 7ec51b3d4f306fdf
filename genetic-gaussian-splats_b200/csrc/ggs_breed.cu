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
// Randomness is counter based (Philox4x32-10 keyed by the seed, counters = generation, child,
// splat, stream), so a step is reproducible for a given (seed, generation) and independent of
// the launch geometry.  The operators draw from the same distributions as the reference; the
// random streams necessarily differ (GA trajectory parity is not a goal).
#include <math.h>

#include "ggs_common.cuh"

namespace ggs {
namespace {

constexpr int kBreedThreads = 256;

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

__device__ __forceinline__ float u01(unsigned v) { return (float)(v >> 8) * (1.0f / 16777216.0f); }

// Two standard normals from two 32-bit words (Box-Muller).
__device__ __forceinline__ float2 normal2(unsigned a, unsigned b)
{
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    return make_float2(r * c, r * s);
}

enum Stream : unsigned { kSelect = 1, kCross = 2, kMask = 3, kNoise0 = 4, kNoise1 = 5, kNoise2 = 6,
                         kForce = 7, kSwapIdx = 8, kSwapScore = 9 };

struct BreedParams {
    const float *pop;
    const float *fitness;
    float *off;
    int P, N, cols;
    int tour_k;
    float cxpb, mutpb;
    float s_xy, s_alog, s_blog, s_theta, s_rgb, s_alpha;
    float log_lo, log_hi;
    unsigned seed_lo, seed_hi, gen;
};

// Bernoulli(mutpb) masks of one splat: bit0,1 = x,y  bit2,3 = log sx, log sy  bit4 = theta
// bit5 = rgb flag  bit6 = alpha flag   (genetic.py:42-50)
__device__ __forceinline__ unsigned gene_masks(const BreedParams &q, unsigned child, unsigned n)
{
    const U4 a = philox4x32_10(q.gen, child, n, kMask, q.seed_lo, q.seed_hi);
    const U4 b = philox4x32_10(q.gen, child, n, kMask + 64u, q.seed_lo, q.seed_hi);
    unsigned m = 0;
    m |= (u01(a.x) < q.mutpb) ? 1u : 0u;
    m |= (u01(a.y) < q.mutpb) ? 2u : 0u;
    m |= (u01(a.z) < q.mutpb) ? 4u : 0u;
    m |= (u01(a.w) < q.mutpb) ? 8u : 0u;
    m |= (u01(b.x) < q.mutpb) ? 16u : 0u;
    m |= (u01(b.y) < q.mutpb) ? 32u : 0u;
    m |= (u01(b.z) < q.mutpb) ? 64u : 0u;
    return m;
}

__device__ __forceinline__ float wrap_angle(float t)
{
    const float kPi = 3.14159265358979323846f, k2Pi = 6.28318530717958647692f;
    float v = fmodf(t + kPi, k2Pi);
    if (v < 0.0f) v += k2Pi;
    return v - kPi;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__global__ void __launch_bounds__(kBreedThreads) breed_kernel(BreedParams q)
{
    __shared__ int s_parent[2];
    __shared__ int s_docx;
    __shared__ int s_count[2][4];   // mutated genes per child and group (xy, ab, theta, colour)
    __shared__ int s_force[2][4];   // forced element when a group came out empty, else -1
    __shared__ unsigned long long s_best[2];
    __shared__ int s_swap_i[2];

    const int tid = threadIdx.x;
    const int pair = blockIdx.x;
    const int nchild = (2 * pair + 1 < q.P) ? 2 : 1;

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
    if (tid < 2) s_best[tid] = 0ull;
    __syncthreads();

    // ---- pass 1: how many genes of each group mutate (for the "at least one" rule)
    for (int c = 0; c < nchild; ++c) {
        const unsigned child = 2 * pair + c;
        int cnt[4] = {0, 0, 0, 0};
        for (int n = tid; n < q.N; n += kBreedThreads) {
            const unsigned m = gene_masks(q, child, n);
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
            if ((tid & 31) == 0 && v) atomicAdd(&s_count[c][g], v);
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

    // ---- pass 2: crossover, mutation, projection; one thread per splat
    const float *pa = q.pop + (int64_t)s_parent[0] * q.N * q.cols;
    const float *pb = q.pop + (int64_t)s_parent[1] * q.N * q.cols;
    for (int n = tid; n < q.N; n += kBreedThreads) {
        bool take_a = true;
        if (s_docx) {
            const U4 r = philox4x32_10(q.gen, pair, n, kCross + 64u, q.seed_lo, q.seed_hi);
            take_a = (r.x & 1u) != 0u;  // p = 0.5 per row (genetic.py:18)
        }
        for (int c = 0; c < nchild; ++c) {
            const unsigned child = 2 * pair + c;
            const float *src = ((c == 0) == take_a ? pa : pb) + (int64_t)n * q.cols;
            float g[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) g[k] = __ldg(src + k);

            unsigned m = gene_masks(q, child, n);
            const int f0 = s_force[c][0], f1 = s_force[c][1], f2 = s_force[c][2], f3 = s_force[c][3];
            if (f0 >= 0 && (f0 >> 1) == n) m |= 1u << (f0 & 1);
            if (f1 >= 0 && (f1 >> 1) == n) m |= 4u << (f1 & 1);
            if (f2 >= 0 && f2 == n) m |= 16u;
            if (f3 >= 0 && (f3 >> 1) == n) m |= 32u << (f3 & 1);

            if (m) {
                const U4 r0 = philox4x32_10(q.gen, child, n, kNoise0, q.seed_lo, q.seed_hi);
                const U4 r1 = philox4x32_10(q.gen, child, n, kNoise1, q.seed_lo, q.seed_hi);
                const U4 r2 = philox4x32_10(q.gen, child, n, kNoise2, q.seed_lo, q.seed_hi);
                const float2 z01 = normal2(r0.x, r0.y), z23 = normal2(r0.z, r0.w);
                const float2 z45 = normal2(r1.x, r1.y), z67 = normal2(r1.z, r1.w);
                const float2 z89 = normal2(r2.x, r2.y);
                if (m & 1u) g[0] += z01.x * q.s_xy;
                if (m & 2u) g[1] += z01.y * q.s_xy;
                if (m & 4u) g[2] += z23.x * q.s_alog;
                if (m & 8u) g[3] += z23.y * q.s_blog;
                if (m & 16u) g[4] += z45.x * q.s_theta;
                if (m & 32u) {  // one flag for the three colour channels, independent noise
                    g[5] += z45.y * q.s_rgb;
                    g[6] += z67.x * q.s_rgb;
                    g[7] += z67.y * q.s_rgb;
                }
                if (m & 64u) g[8] += z89.x * q.s_alpha;
            }
            // projection onto the legal box (utils.py:36-45)
            g[0] = clampf(g[0], 0.0f, 1.0f);
            g[1] = clampf(g[1], 0.0f, 1.0f);
            g[2] = clampf(g[2], q.log_lo, q.log_hi);
            g[3] = clampf(g[3], q.log_lo, q.log_hi);
            g[4] = wrap_angle(g[4]);
#pragma unroll
            for (int k = 5; k < 9; ++k) g[k] = clampf(g[k], 0.0f, 255.0f);

            float *dst = q.off + ((int64_t)child * q.N + n) * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) dst[k] = g[k];
        }
    }
    __syncthreads();  // the child rows written above are read back below by other threads

    // ---- pass 3: bring a bigger later splat forward (genetic.py:74-91)
    if (q.N >= 2) {
        if (tid < nchild) {
            const U4 r = philox4x32_10(q.gen, 2 * pair + tid, 0, kSwapIdx, q.seed_lo, q.seed_hi);
            s_swap_i[tid] = (int)__umulhi(r.x, (unsigned)(q.N - 1));  // uniform in [0, N-2]
        }
        __syncthreads();
        for (int c = 0; c < nchild; ++c) {
            const unsigned child = 2 * pair + c;
            const float *row = q.off + (int64_t)child * q.N * 9;
            const int i = s_swap_i[c];
            const float size_i = row[i * 9 + 2] + row[i * 9 + 3];  // log(sigma_x * sigma_y)
            unsigned long long best = 0ull;
            for (int j = i + 1 + tid; j < q.N; j += kBreedThreads) {
                if (row[j * 9 + 2] + row[j * 9 + 3] > size_i) {
                    // uniform choice among the candidates = arg max of iid scores
                    const U4 r = philox4x32_10(q.gen, child, j, kSwapScore, q.seed_lo, q.seed_hi);
                    const unsigned long long key = ((unsigned long long)(r.x | 1u) << 32) | (unsigned)j;
                    best = key > best ? key : best;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if ((tid & 31) == 0 && best) atomicMax(&s_best[c], best);
        }
        __syncthreads();
        if (tid < 9 * nchild) {
            const int c = tid / 9, k = tid - 9 * c;
            if (s_best[c] != 0ull) {
                float *row = q.off + (int64_t)(2 * pair + c) * q.N * 9;
                const int i = s_swap_i[c], j = (int)(s_best[c] & 0xffffffffull);
                const float a = row[i * 9 + k], b = row[j * 9 + k];
                row[i * 9 + k] = b;
                row[j * 9 + k] = a;
            }
        }
    }
}

}  // namespace

cudaError_t launch_breed(const float *d_pop, const float *d_fitness, int P, int N, int cols,
                         float *d_offspring, int tour_k, float cxpb, float mutpb,
                         const float sigma6[6], float log_lo, float log_hi, uint64_t seed,
                         uint32_t generation, cudaStream_t stream)
{
    if (P <= 0 || N <= 0) return cudaSuccess;
    BreedParams q;
    q.pop = d_pop;
    q.fitness = d_fitness;
    q.off = d_offspring;
    q.P = P;
    q.N = N;
    q.cols = cols;
    q.tour_k = tour_k;
    q.cxpb = cxpb;
    q.mutpb = mutpb;
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
    breed_kernel<<<(P + 1) / 2, kBreedThreads, 0, stream>>>(q);
    return cudaGetLastError();
}

}  // namespace ggs
