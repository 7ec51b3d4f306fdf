// GA generations without the host in the loop (SURVEY.md section 8f "next" #1: the batched
// on-device GA step *including elitism*, algorithm.py:87-160).
//
// The reference's default run is 500,000 generations of a 32-individual population
// (modules/config.py): each generation is tens of microseconds of GPU work, so what decides the
// run time is whether the host has to look at the fitness vector in between.  Here it does not:
// a generation is four launches on one stream --
//     breed  (ggs_breed.cu)   parents + fitness  -> children, written behind the elite rows
//     decode + raster         children           -> their fitness
//     select (below)          elitism, the new fitness vector, its stable ranking, the
//                             (best, mean, median) curve point, best-so-far bookkeeping
// -- and ggs_ga_run() enqueues as many generations as the caller asks for.  The host reads the
// curves and the best individual when it wants them (ggs_ga_state), e.g. once per video frame.
//
// State lives in two generation buffers of [n_elite + P][N][9] floats used alternately: the
// population of generation g is rows [0, P) of buffer g & 1 (elites first, then the first
// P - n_elite children, algorithm.py:128-141).
#include <math.h>

#include <new>

#include "ggs_common.cuh"

namespace ggs {
namespace {

constexpr int kSelectThreads = 1024;
constexpr int kMaxEnginePop = 16384;  // ranking sorts the population in one CTA's shared memory

__device__ __forceinline__ unsigned sortable(float v)
{
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Copies one genome row with all of a thread's loads in flight before its first store (16-byte
// accesses when both ends allow it): one memory round trip per 4 x blockDim float4s instead of
// one per element.
__device__ __forceinline__ void copy_row(float *__restrict__ to, const float *__restrict__ from,
                                         int64_t n, int tid, int nthreads)
{
    int64_t done = 0;
    if (((reinterpret_cast<uintptr_t>(to) | reinterpret_cast<uintptr_t>(from)) & 15u) == 0) {
        const int64_t n4 = n >> 2;
        const float4 *f4 = reinterpret_cast<const float4 *>(from);
        float4 *t4 = reinterpret_cast<float4 *>(to);
        for (int64_t base = 0; base < n4; base += 4 * (int64_t)nthreads) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = base + (int64_t)u * nthreads + tid;
                if (i < n4) v[u] = f4[i];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = base + (int64_t)u * nthreads + tid;
                if (i < n4) t4[i] = v[u];
            }
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) to[i] = from[i];
}

struct SelectParams {
    const float *pop;        // current population  [P][N][9]   (rows [0,P) of its buffer)
    const float *fit;        // its fitness         [P]
    const int *order_in;     // stable ascending ranking of `fit`
    int *order_out;          // ranking of new_fit (a different buffer: the copy CTAs read order_in)
    float *next;             // next buffer [n_elite + keep ...][N][9]; children already at row n_elite
    const float *child_fit;  // [keep]
    float *new_fit;          // [P]
    double *curve;           // this generation's (best, mean, median)
    float *best_ind;         // [N][9]
    double *best_fit;        // in / out
    int *no_improve;         // in / out
    int P, N, n_elite;
    // sharded evaluation: child_fit is this rank's gathered vector; wait for every rank's values
    const unsigned *wait_flags;
    int *wait_status;
    int wait_world;
    unsigned wait_epoch;
};

// Grid = 1 + n_elite CTAs.  CTA e + 1 copies elite e (the e-th best of the current generation,
// algorithm.py:128-141) to the front of the next population; CTA 0 assembles the new fitness
// vector, ranks it and keeps the statistics (algorithm.py:143-160).
__global__ void __launch_bounds__(kSelectThreads) select_kernel(SelectParams q)
{
    extern __shared__ __align__(16) unsigned long long s_key[];  // [P2] fitness key << 32 | index
    __shared__ double s_sum[kSelectThreads / 32];
    __shared__ int s_improved;
    const int tid = threadIdx.x;
    const int keep = q.P - q.n_elite;
    const int64_t row = (int64_t)q.N * 9;
    pdl_wait();
    pdl_trigger();

    if (blockIdx.x > 0) {  // elites survive unchanged, in rank order
        const int e = blockIdx.x - 1;
        copy_row(q.next + e * row, q.pop + q.order_in[e] * row, row, tid, kSelectThreads);
        return;
    }
    if (q.wait_world > 0) {  // the children's fitness arrives from the other GPUs (ggs_peers.cu)
        peer_wait(q.wait_flags, q.wait_world, q.wait_epoch, q.wait_status);
        __syncthreads();
    }
    for (int e = tid; e < q.n_elite; e += kSelectThreads) q.new_fit[e] = q.fit[q.order_in[e]];
    for (int i = tid; i < keep; i += kSelectThreads) q.new_fit[q.n_elite + i] = __ldcv(q.child_fit + i);
    __syncthreads();

    // stable ascending ranking: the index in the low word makes every key distinct
    int P2 = 1;
    while (P2 < q.P) P2 <<= 1;
    for (int i = tid; i < P2; i += kSelectThreads)
        s_key[i] = (i < q.P) ? (((unsigned long long)sortable(q.new_fit[i]) << 32) | (unsigned)i)
                             : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += kSelectThreads) {
                const int partner = i ^ j;
                if (partner > i) {
                    const unsigned long long a = s_key[i], b = s_key[partner];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        s_key[i] = b;
                        s_key[partner] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    double part = 0.0;
    for (int i = tid; i < q.P; i += kSelectThreads) {
        const int idx = (int)(s_key[i] & 0xffffffffull);
        q.order_out[i] = idx;
        part += (double)q.new_fit[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) s_sum[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double sum = 0.0;
        for (int w = 0; w < kSelectThreads / 32; ++w) sum += s_sum[w];
        const double best = (double)q.new_fit[(int)(s_key[0] & 0xffffffffull)];
        const double lo = (double)q.new_fit[(int)(s_key[(q.P - 1) / 2] & 0xffffffffull)];
        const double hi = (double)q.new_fit[(int)(s_key[q.P / 2] & 0xffffffffull)];
        q.curve[1] = sum / (double)q.P;
        q.curve[2] = 0.5 * (lo + hi);  // statistics.median: mean of the middle two
        // best so far (algorithm.py:150-155)
        const bool improved = best + 1e-10 < *q.best_fit;
        if (improved) {
            *q.best_fit = best;
            *q.no_improve = 0;
        } else {
            *q.no_improve += 1;
        }
        q.curve[0] = *q.best_fit;
        s_improved = improved ? 1 : 0;
    }
    __syncthreads();
    if (s_improved) {
        // an elite row of `next` may still be in flight in its copy CTA: read its source instead
        const int idx = (int)(s_key[0] & 0xffffffffull);
        const float *from = (idx < q.n_elite) ? q.pop + q.order_in[idx] * row : q.next + idx * row;
        copy_row(q.best_ind, from, row, tid, kSelectThreads);
    }
}

// ---- simulated annealing: the Metropolis step of one iteration (annealing.py:125-137) ----------
struct MetropolisParams {
    const float *cand;     // [tries][N][9] neighbours, all proposed from the current state
    const float *energy;   // [tries]
    float *current;        // [N][9]   in / out
    float *best;           // [N][9]   in / out
    double *e_current;     // in / out
    double *e_best;        // in / out
    double *curve;         // this iteration's (best, current)
    double temperature;
    double uniform[kMaxTries];  // one U[0,1) draw per try (used when the try is uphill)
    int tries, N;
};

// One CTA.  Thread 0 applies the tries in order: accept when dE <= 0 or u < exp(-dE / T); the
// best-so-far test follows every try as in the reference.  Then all threads move the rows.
__global__ void __launch_bounds__(kSelectThreads) metropolis_kernel(MetropolisParams q)
{
    __shared__ int s_cur, s_best;
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0) {
        double e_cur = *q.e_current, e_best = *q.e_best;
        int cur = -1, best = -1;
        for (int k = 0; k < q.tries; ++k) {
            const double e_new = (double)q.energy[k];
            const double dE = e_new - e_cur;
            if (dE <= 0.0 || (q.temperature > 0.0 && q.uniform[k] < exp(-dE / q.temperature))) {
                cur = k;
                e_cur = e_new;
            }
            if (e_cur + 1e-12 < e_best) {
                e_best = e_cur;
                best = cur;
            }
        }
        *q.e_current = e_cur;
        *q.e_best = e_best;
        q.curve[0] = e_best;
        q.curve[1] = e_cur;
        s_cur = cur;
        s_best = best;
    }
    __syncthreads();
    const int64_t row = (int64_t)q.N * 9;
    if (s_best >= 0) copy_row(q.best, q.cand + s_best * row, row, threadIdx.x, kSelectThreads);
    if (s_cur >= 0) copy_row(q.current, q.cand + s_cur * row, row, threadIdx.x, kSelectThreads);
}

// The engines own their workspace, cleared once at creation: the ticket counters at its head are
// zero before every evaluation, so the fused raster needs no memset in front of it.
EvalOptions owned_workspace()
{
    EvalOptions o;
    o.counters_zeroed = true;
    return o;
}

int fail(cudaError_t e, const char *what)
{
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GGS_ECUDA;
}
#define GGS_TRY(call)                                    \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return fail(e_, #call);   \
    } while (0)

}  // namespace
}  // namespace ggs

using namespace ggs;

struct ggs_ga {
    int device = 0;
    int P = 0, N = 0, H = 0, W = 0, n_elite = 0, capacity = 0;  // capacity: curve points
    float *room[2] = {nullptr, nullptr};  // generation buffers [n_elite + P][N][9]
    float *fit[2] = {nullptr, nullptr};   // [P]
    float *child_fit = nullptr;           // [P]
    int *order[2] = {nullptr, nullptr};   // [P] rankings, used alternately
    int ord = 0;                          // which one ranks the current population
    float *target = nullptr, *mask = nullptr;
    int mode = GGS_MODE_PLAIN;
    float beta = 1.0f, k_sigma = 3.0f;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    double *curves = nullptr;             // [capacity][3]
    float *best_ind = nullptr;            // [N][9]
    double *best_fit = nullptr;
    int *no_improve = nullptr;
    uint64_t seed = 0;
    int cur = 0;          // buffer holding the current population
    int generation = -1;  // generations completed; -1 until ggs_ga_start
    bool has_target = false;
    ggs_peers *peers = nullptr;  // evaluation sharded over the ranks of this group, or NULL
    size_t select_smem_granted = 0;
};

// Contiguous split of `total` items; the first total % world ranks get one more (the rule of
// ggs_b200.distributed.shard_bounds).
static void shard_bounds(int total, int world, int rank, int *lo, int *hi)
{
    const int base = total / world, extra = total % world;
    *lo = rank * base + (rank < extra ? rank : extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
}

// Fitness of `count` individuals starting at `first` -> fit_out[count] (g->child_fit on one GPU).
// Sharded: this rank evaluates its slice with the configuration of the whole batch and every
// rank's values land in the gathered vector of a new epoch, returned through *fit_out / *epoch;
// the consumer (select_kernel) waits for the flags.
static int ga_evaluate(ggs_ga *g, const float *first, int count, const float **fit_out, unsigned *epoch,
                       cudaStream_t st)
{
    const float bg[3] = {1.0f, 1.0f, 1.0f};
    const float *mask = g->mode == GGS_MODE_PLAIN ? nullptr : g->mask;
    EvalOptions opt = owned_workspace();
    *epoch = 0;
    if (g->peers == nullptr) {
        *fit_out = g->child_fit;
        return evaluate(first, GGS_LAYOUT_AXES_ANGLE, count, g->N, 9, g->H, g->W, g->k_sigma, bg, g->target,
                        mask, g->mode, g->beta, g->child_fit, nullptr, 0, g->ws, g->ws_bytes, st, opt);
    }
    int lo = 0, hi = 0;
    shard_bounds(count, peers_world(g->peers), peers_rank(g->peers), &lo, &hi);
    opt.split = choose_split(count, g->N, g->H, g->W);
    opt.peers = peers_next(g->peers, lo);
    *epoch = opt.peers.epoch;
    float *mine = peers_gathered(g->peers, opt.peers.epoch);
    *fit_out = mine;
    if (hi > lo)
        return evaluate(first + (size_t)lo * g->N * 9, GGS_LAYOUT_AXES_ANGLE, hi - lo, g->N, 9, g->H, g->W,
                        g->k_sigma, bg, g->target, mask, g->mode, g->beta, mine + lo, nullptr, 0, g->ws,
                        g->ws_bytes, st, opt);
    GGS_TRY(peers_signal_empty(opt.peers, st));
    return GGS_OK;
}

static int ga_rank(ggs_ga *g, int n_elite, int next_buf, const float *child_fit, unsigned epoch,
                   cudaStream_t st)
{
    SelectParams q;
    q.pop = g->room[g->cur];
    q.fit = g->fit[g->cur];
    q.order_in = g->order[g->ord];
    q.order_out = g->order[g->ord ^ 1];
    q.next = g->room[next_buf];
    q.child_fit = child_fit;
    q.new_fit = g->fit[next_buf];
    q.wait_world = (g->peers != nullptr && epoch != 0) ? peers_world(g->peers) : 0;
    q.wait_flags = q.wait_world ? peers_flags(g->peers) : nullptr;
    q.wait_status = q.wait_world ? peers_status(g->peers) : nullptr;
    q.wait_epoch = epoch;
    q.curve = g->curves + (size_t)(g->generation + 1) * 3;
    q.best_ind = g->best_ind;
    q.best_fit = g->best_fit;
    q.no_improve = g->no_improve;
    q.P = g->P;
    q.N = g->N;
    q.n_elite = n_elite;
    int P2 = 1;
    while (P2 < g->P) P2 <<= 1;
    const size_t smem = (size_t)P2 * sizeof(unsigned long long);
    if (smem > 40 * 1024 && g->select_smem_granted < smem) {  // once per engine, not per generation
        GGS_TRY(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        g->select_smem_granted = smem;
    }
    GGS_TRY(launch_kernel(select_kernel, 1 + n_elite, kSelectThreads, smem, st, q));
    g->ord ^= 1;
    return GGS_OK;
}

extern "C" {

int ggs_ga_create(int device, int P, int N, int H, int W, int n_elite, int max_generations,
                  ggs_ga **out)
{
    if (out == nullptr) {
        set_error("ggs_ga_create: out is NULL");
        return GGS_EINVAL;
    }
    *out = nullptr;
    if (P < 1 || P > kMaxEnginePop || N < 1 || H < 1 || W < 1 || H > GGS_MAX_SIDE || W > GGS_MAX_SIDE ||
        n_elite < 0 || n_elite > P || max_generations < 0) {
        set_error("ggs_ga_create: bad arguments (P=%d, at most %d; N=%d; %dx%d; n_elite=%d; "
                  "max_generations=%d)", P, kMaxEnginePop, N, H, W, n_elite, max_generations);
        return GGS_EINVAL;
    }
    int n = 0;
    GGS_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error("ggs_ga_create: device %d not visible (%d devices)", device, n);
        return GGS_ENODEVICE;
    }
    DeviceGuard on_device(device);
    GGS_TRY(on_device.status());
    ggs_ga *g = new (std::nothrow) ggs_ga();
    if (!g) {
        set_error("out of host memory");
        return GGS_EINVAL;
    }
    g->device = device;
    g->P = P;
    g->N = N;
    g->H = H;
    g->W = W;
    g->n_elite = n_elite;
    g->capacity = max_generations + 1;
    const size_t rows = (size_t)(n_elite + P) * N * 9 * sizeof(float);
    g->ws_bytes = workspace_bytes(P, N, H, W);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc(&g->room[i], rows);
        if (e == cudaSuccess) e = cudaMalloc(&g->fit[i], (size_t)P * sizeof(float));
    }
    if (e == cudaSuccess) e = cudaMalloc(&g->child_fit, (size_t)P * sizeof(float));
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&g->order[i], (size_t)P * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&g->target, (size_t)H * W * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->mask, (size_t)H * W * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->ws, g->ws_bytes);
    if (e == cudaSuccess) e = cudaMemset(g->ws, 0, g->ws_bytes);  // ticket counters start at zero
    if (e == cudaSuccess) e = cudaMalloc(&g->curves, (size_t)g->capacity * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&g->best_ind, (size_t)N * 9 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->best_fit, sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&g->no_improve, sizeof(int));
    if (e != cudaSuccess) {
        ggs_ga_destroy(g);
        return fail(e, "ggs_ga_create: cudaMalloc");
    }
    *out = g;
    return GGS_OK;
}

void ggs_ga_destroy(ggs_ga *g)
{
    if (!g) return;
    DeviceGuard on_device(g->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) {
        if (g->room[i]) cudaFree(g->room[i]);
        if (g->fit[i]) cudaFree(g->fit[i]);
    }
    void *rest[] = {g->child_fit, g->order[0], g->order[1], g->target, g->mask, g->ws, g->curves, g->best_ind,
                    g->best_fit, g->no_improve};
    for (void *p : rest)
        if (p) cudaFree(p);
    delete g;
}

int ggs_ga_set_target(ggs_ga *g, const float *d_target, const float *d_mask, int mode,
                      float boost_beta, float k_sigma, void *stream)
{
    if (!g || !d_target) {
        set_error("ggs_ga_set_target: NULL argument");
        return GGS_EINVAL;
    }
    if (mode < GGS_MODE_PLAIN || mode > GGS_MODE_BOOST || (mode != GGS_MODE_PLAIN && !d_mask)) {
        set_error("ggs_ga_set_target: mode %d needs a mask", mode);
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    GGS_TRY(cudaMemcpyAsync(g->target, d_target, (size_t)g->H * g->W * 3 * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
    if (d_mask)
        GGS_TRY(cudaMemcpyAsync(g->mask, d_mask, (size_t)g->H * g->W * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    g->mode = mode;
    g->beta = boost_beta;
    g->k_sigma = k_sigma;
    g->has_target = true;
    return GGS_OK;
}

int ggs_ga_start(ggs_ga *g, const float *d_population, int cols, uint64_t seed, void *stream)
{
    if (!g || !d_population || cols < 9) {
        set_error("ggs_ga_start: bad arguments");
        return GGS_EINVAL;
    }
    if (!g->has_target) {
        set_error("ggs_ga_start: call ggs_ga_set_target first");
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    // generation 0: the given population (first 9 columns), evaluated and ranked
    GGS_TRY(cudaMemcpy2DAsync(g->room[0], 9 * sizeof(float), d_population, (size_t)cols * sizeof(float),
                              9 * sizeof(float), (size_t)g->P * g->N, cudaMemcpyDeviceToDevice, st));
    const float *fit = nullptr;
    unsigned epoch = 0;
    int rc = ga_evaluate(g, g->room[0], g->P, &fit, &epoch, st);
    if (rc) return rc;
    const double inf = INFINITY;
    const int zero = 0;
    GGS_TRY(cudaMemcpyAsync(g->best_fit, &inf, sizeof(double), cudaMemcpyHostToDevice, st));
    GGS_TRY(cudaMemcpyAsync(g->no_improve, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    g->seed = seed;
    g->cur = 0;
    g->generation = -1;
    rc = ga_rank(g, 0, 0, fit, epoch, st);  // no elites: new_fit = child_fit, rows already in place
    if (rc) return rc;
    g->generation = 0;
    return GGS_OK;
}

int ggs_ga_run(ggs_ga *g, int count, const float *h_sigma6, int tour_k, float cxpb, float mutpb,
               float log_scale_lo, float log_scale_hi, void *stream)
{
    if (!g || count < 0 || (count > 0 && !h_sigma6) || tour_k < 1) {
        set_error("ggs_ga_run: bad arguments");
        return GGS_EINVAL;
    }
    if (g->generation < 0) {
        set_error("ggs_ga_run: call ggs_ga_start first");
        return GGS_EINVAL;
    }
    if (g->generation + count >= g->capacity) {
        set_error("ggs_ga_run: %d more generations exceed the %d the engine was created for",
                  count, g->capacity - 1);
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    const int keep = g->P - g->n_elite;
    const size_t row = (size_t)g->N * 9;
    for (int k = 0; k < count; ++k) {
        const int gen = g->generation + 1, nb = g->cur ^ 1;
        float *children = g->room[nb] + (size_t)g->n_elite * row;
        GGS_TRY(launch_breed(g->room[g->cur], g->fit[g->cur], g->P, g->N, 9, keep, children, tour_k,
                             cxpb, mutpb, h_sigma6 + 6 * (size_t)k, log_scale_lo, log_scale_hi,
                             g->seed, (uint32_t)gen, st));
        const float *fit = nullptr;
        unsigned epoch = 0;
        int rc = ga_evaluate(g, children, keep, &fit, &epoch, st);
        if (rc) return rc;
        rc = ga_rank(g, g->n_elite, nb, fit, epoch, st);
        if (rc) return rc;
        g->cur = nb;
        g->generation = gen;
    }
    return GGS_OK;
}

int ggs_ga_state(ggs_ga *g, void *stream, int *h_generation, double *h_best_fitness,
                 int *h_no_improve, double *h_curves3, int curves_from, float *h_best_individual)
{
    if (!g || g->generation < 0) {
        set_error("ggs_ga_state: engine not started");
        return GGS_EINVAL;
    }
    if (curves_from < 0 || curves_from > g->generation + 1) {
        set_error("ggs_ga_state: curves_from %d outside [0, %d]", curves_from, g->generation + 1);
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    if (h_best_fitness)
        GGS_TRY(cudaMemcpyAsync(h_best_fitness, g->best_fit, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (h_no_improve)
        GGS_TRY(cudaMemcpyAsync(h_no_improve, g->no_improve, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (h_curves3 && curves_from <= g->generation)
        GGS_TRY(cudaMemcpyAsync(h_curves3, g->curves + (size_t)curves_from * 3,
                                (size_t)(g->generation + 1 - curves_from) * 3 * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    if (h_best_individual)
        GGS_TRY(cudaMemcpyAsync(h_best_individual, g->best_ind, (size_t)g->N * 9 * sizeof(float),
                                cudaMemcpyDeviceToHost, st));
    GGS_TRY(cudaStreamSynchronize(st));
    if (h_generation) *h_generation = g->generation;
    return GGS_OK;
}

int ggs_ga_set_peers(ggs_ga *g, ggs_peers *peers)
{
    if (!g) {
        set_error("ggs_ga_set_peers: NULL engine");
        return GGS_EINVAL;
    }
    if (g->generation >= 0) {
        set_error("ggs_ga_set_peers: call before ggs_ga_start");
        return GGS_EINVAL;
    }
    if (peers != nullptr && (!peers_ready(peers) || peers_capacity(peers) < g->P)) {
        set_error("ggs_ga_set_peers: peers not connected, or their capacity is below the population (%d)", g->P);
        return GGS_EINVAL;
    }
    g->peers = (peers != nullptr && peers_world(peers) > 1) ? peers : nullptr;
    return GGS_OK;
}

int ggs_ga_population(ggs_ga *g, const float **d_population, const float **d_fitness)
{
    if (!g || g->generation < 0) {
        set_error("ggs_ga_population: engine not started");
        return GGS_EINVAL;
    }
    if (d_population) *d_population = g->room[g->cur];
    if (d_fitness) *d_fitness = g->fit[g->cur];
    return GGS_OK;
}

/* ---------------------------------------------------------------------------------------- */
/* simulated annealing                                                                        */

}  // extern "C"

struct ggs_sa {
    int device = 0;
    int N = 0, H = 0, W = 0, tries = 0, capacity = 0;
    float *current = nullptr, *best = nullptr;  // [N][9]
    float *cand2[2] = {nullptr, nullptr};        // [tries][N][9] each: proposals alternate, so that a
                                                 // proposal can read an accepted candidate of the
                                                 // previous batch while it writes the next one
    int cw = 0;                                  // the buffer holding the latest proposals
    float *energy = nullptr;                     // [tries]
    float *dummy_fit = nullptr;                  // [1] fitness of the single "parent"
    float *target = nullptr, *mask = nullptr;
    int mode = GGS_MODE_PLAIN;
    float beta = 1.0f, k_sigma = 3.0f;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    double *curves = nullptr;  // [capacity][2]
    double *e_state = nullptr;  // [2][2]: {e_current, e_best}, two copies used alternately (a judging
                                // launch reads one and writes the other)
    int ep = 0;                 // the copy that holds the settled values
    uint64_t seed = 0;
    int iteration = -1;
    bool has_target = false;
    bool sequential = true;  // the reference's chain (annealing.py:121-146); false: batched neighbours
    // the evaluation still waiting for its Metropolis step: the next proposal launch judges it
    struct Pending {
        bool on = false;
        int tries = 0, cand = 0;
        double temperature = 0.0;
        double uniform[kMaxTries] = {};
        double *curve = nullptr;
    } pending;
};

extern "C" {

int ggs_sa_create(int device, int N, int H, int W, int tries, int max_iterations, ggs_sa **out)
{
    if (out == nullptr) {
        set_error("ggs_sa_create: out is NULL");
        return GGS_EINVAL;
    }
    *out = nullptr;
    if (N < 1 || H < 1 || W < 1 || H > GGS_MAX_SIDE || W > GGS_MAX_SIDE || tries < 1 ||
        tries > kMaxTries || max_iterations < 0) {
        set_error("ggs_sa_create: bad arguments (N=%d; %dx%d; tries=%d, at most %d; "
                  "max_iterations=%d)", N, H, W, tries, kMaxTries, max_iterations);
        return GGS_EINVAL;
    }
    int n = 0;
    GGS_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error("ggs_sa_create: device %d not visible (%d devices)", device, n);
        return GGS_ENODEVICE;
    }
    DeviceGuard on_device(device);
    GGS_TRY(on_device.status());
    ggs_sa *g = new (std::nothrow) ggs_sa();
    if (!g) {
        set_error("out of host memory");
        return GGS_EINVAL;
    }
    g->device = device;
    g->N = N;
    g->H = H;
    g->W = W;
    g->tries = tries;
    g->capacity = max_iterations + 1;
    const size_t row = (size_t)N * 9 * sizeof(float);
    g->ws_bytes = workspace_bytes(tries, N, H, W);
    cudaError_t e = cudaMalloc(&g->current, row);
    if (e == cudaSuccess) e = cudaMalloc(&g->best, row);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&g->cand2[i], row * tries);
    if (e == cudaSuccess) e = cudaMalloc(&g->energy, (size_t)tries * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->dummy_fit, sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(g->dummy_fit, 0, sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->target, (size_t)H * W * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->mask, (size_t)H * W * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&g->ws, g->ws_bytes);
    if (e == cudaSuccess) e = cudaMemset(g->ws, 0, g->ws_bytes);  // ticket counters start at zero
    if (e == cudaSuccess) e = cudaMalloc(&g->curves, (size_t)g->capacity * 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&g->e_state, 4 * sizeof(double));
    if (e != cudaSuccess) {
        ggs_sa_destroy(g);
        return fail(e, "ggs_sa_create: cudaMalloc");
    }
    *out = g;
    return GGS_OK;
}

void ggs_sa_destroy(ggs_sa *g)
{
    if (!g) return;
    DeviceGuard on_device(g->device);
    cudaDeviceSynchronize();
    void *all[] = {g->current, g->best, g->cand2[0], g->cand2[1], g->energy, g->dummy_fit, g->target,
                   g->mask, g->ws, g->curves, g->e_state};
    for (void *p : all)
        if (p) cudaFree(p);
    delete g;
}

int ggs_sa_set_target(ggs_sa *g, const float *d_target, const float *d_mask, int mode,
                      float boost_beta, float k_sigma, void *stream)
{
    if (!g || !d_target) {
        set_error("ggs_sa_set_target: NULL argument");
        return GGS_EINVAL;
    }
    if (mode < GGS_MODE_PLAIN || mode > GGS_MODE_BOOST || (mode != GGS_MODE_PLAIN && !d_mask)) {
        set_error("ggs_sa_set_target: mode %d needs a mask", mode);
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    GGS_TRY(cudaMemcpyAsync(g->target, d_target, (size_t)g->H * g->W * 3 * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
    if (d_mask)
        GGS_TRY(cudaMemcpyAsync(g->mask, d_mask, (size_t)g->H * g->W * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    g->mode = mode;
    g->beta = boost_beta;
    g->k_sigma = k_sigma;
    g->has_target = true;
    return GGS_OK;
}

static int sa_energy(ggs_sa *g, const float *genomes, int B, cudaStream_t st, bool decoded = false)
{
    const float bg[3] = {1.0f, 1.0f, 1.0f};
    EvalOptions opt = owned_workspace();
    opt.decoded = decoded;  // the proposal kernel has written the records of these B candidates
    return evaluate(genomes, GGS_LAYOUT_AXES_ANGLE, B, g->N, 9, g->H, g->W, g->k_sigma, bg, g->target,
                    g->mode == GGS_MODE_PLAIN ? nullptr : g->mask, g->mode, g->beta, g->energy,
                    nullptr, 0, g->ws, g->ws_bytes, st, opt);
}

// The Metropolis step the pending evaluation is waiting for, as the proposal kernel's judging stage.
static ProposeJudge sa_judge(ggs_sa *g)
{
    ProposeJudge j;
    if (!g->pending.on) return j;
    j.tries = g->pending.tries;
    j.energy = g->energy;
    j.cand_prev = g->cand2[g->pending.cand];
    j.current = g->current;
    j.best = g->best;
    j.e_in = g->e_state + 2 * g->ep;
    j.e_out = g->e_state + 2 * (g->ep ^ 1);
    j.curve = g->pending.curve;
    j.temperature = g->pending.temperature;
    for (int t = 0; t < g->pending.tries; ++t) j.uniform[t] = g->pending.uniform[t];
    return j;
}

static MetropolisParams sa_metropolis(ggs_sa *g, const float *cand, int tries, double temperature,
                                      const double *uniform, double *curve)
{
    MetropolisParams q = {};
    q.cand = cand;
    q.energy = g->energy;
    q.current = g->current;
    q.best = g->best;
    q.e_current = g->e_state + 2 * g->ep;
    q.e_best = g->e_state + 2 * g->ep + 1;
    q.curve = curve;
    q.temperature = temperature;
    for (int t = 0; t < tries; ++t) q.uniform[t] = uniform[t];
    q.tries = tries;
    q.N = g->N;
    return q;
}

// One batch of `count` proposals (mutated copies of the state the chain is in), evaluated, with the
// Metropolis draws (temperature, uniform[count]) that will judge them and the curve point they
// belong to.  With a state that fits the proposal kernel (ggs_breed.cu) this is TWO launches:
// the proposal launch first judges the previous batch, then mutates and decodes, and the raster
// scores the children; their own judgement is left pending for the next proposal (or for
// sa_settle).  Larger states take breed + decode + raster + Metropolis.
static int sa_batch(ggs_sa *g, int count, const float *sigma6, float mutpb, float log_lo, float log_hi,
                    uint32_t number, double temperature, const double *uniform, double *curve,
                    cudaStream_t st)
{
    if (propose_possible(g->N, 9)) {
        const ProposeJudge judge = sa_judge(g);
        const int w = g->cw ^ 1;
        GGS_TRY(launch_propose(g->current, g->N, 9, count, g->cand2[w], mutpb, sigma6, log_lo, log_hi,
                               g->seed, number, carve_workspace(g->ws, count, g->N, g->H, g->W), g->H,
                               g->W, g->k_sigma, judge, st));
        if (judge.tries > 0) g->ep ^= 1;
        g->cw = w;
        int rc = sa_energy(g, g->cand2[w], count, st, /*decoded=*/true);
        if (rc) return rc;
        g->pending.on = true;
        g->pending.tries = count;
        g->pending.cand = w;
        g->pending.temperature = temperature;
        for (int t = 0; t < count; ++t) g->pending.uniform[t] = uniform[t];
        g->pending.curve = curve;
        return GGS_OK;
    }
    GGS_TRY(launch_breed(g->current, g->dummy_fit, 1, g->N, 9, count, g->cand2[0], 1, 0.0f, mutpb, sigma6,
                         log_lo, log_hi, g->seed, number, st));
    int rc = sa_energy(g, g->cand2[0], count, st);
    if (rc) return rc;
    GGS_TRY(launch_kernel(metropolis_kernel, 1, kSelectThreads, 0, st,
                          sa_metropolis(g, g->cand2[0], count, temperature, uniform, curve)));
    return GGS_OK;
}

// Judge whatever is still pending, so that current / best / energies / curves are final.
static int sa_settle(ggs_sa *g, cudaStream_t st)
{
    if (!g->pending.on) return GGS_OK;
    const float zero6[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    GGS_TRY(launch_propose(g->current, g->N, 9, 0, g->cand2[g->cw ^ 1], 0.0f, zero6, 0.0f, 0.0f, g->seed, 0,
                           carve_workspace(g->ws, 1, g->N, g->H, g->W), g->H, g->W, g->k_sigma,
                           sa_judge(g), st));
    g->ep ^= 1;
    g->pending.on = false;
    return GGS_OK;
}

int ggs_sa_start(ggs_sa *g, const float *d_state, int cols, uint64_t seed, void *stream)
{
    if (!g || !d_state || cols < 9) {
        set_error("ggs_sa_start: bad arguments");
        return GGS_EINVAL;
    }
    if (!g->has_target) {
        set_error("ggs_sa_start: call ggs_sa_set_target first");
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    // iteration 0: the given state is the current and the best one; its energy opens the curves.
    // It is staged as candidate 0 and "accepted" by the Metropolis kernel (dE = -inf).
    g->cw = 0;
    g->ep = 0;
    g->pending.on = false;
    GGS_TRY(cudaMemcpy2DAsync(g->cand2[0], 9 * sizeof(float), d_state, (size_t)cols * sizeof(float),
                              9 * sizeof(float), (size_t)g->N, cudaMemcpyDeviceToDevice, st));
    int rc = sa_energy(g, g->cand2[0], 1, st);
    if (rc) return rc;
    const double inf2[2] = {INFINITY, INFINITY};
    GGS_TRY(cudaMemcpyAsync(g->e_state, inf2, sizeof(inf2), cudaMemcpyHostToDevice, st));
    const double no_draw = 0.0;
    GGS_TRY(launch_kernel(metropolis_kernel, 1, kSelectThreads, 0, st,
                          sa_metropolis(g, g->cand2[0], 1, 0.0, &no_draw, g->curves)));
    g->seed = seed;
    g->iteration = 0;
    return GGS_OK;
}

int ggs_sa_run(ggs_sa *g, int count, const float *h_sigma6, const double *h_temperature,
               const double *h_uniform, float mutpb, float log_scale_lo, float log_scale_hi,
               void *stream)
{
    if (!g || count < 0 || (count > 0 && (!h_sigma6 || !h_temperature || !h_uniform))) {
        set_error("ggs_sa_run: bad arguments");
        return GGS_EINVAL;
    }
    if (g->iteration < 0) {
        set_error("ggs_sa_run: call ggs_sa_start first");
        return GGS_EINVAL;
    }
    if (g->iteration + count >= g->capacity) {
        set_error("ggs_sa_run: %d more iterations exceed the %d the engine was created for", count,
                  g->capacity - 1);
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    for (int k = 0; k < count; ++k) {
        const int it = g->iteration + 1;
        double *curve = g->curves + (size_t)it * 2;
        const float *sigma6 = h_sigma6 + 6 * (size_t)k;
        const double *uniform = h_uniform + (size_t)k * g->tries;
        if (g->sequential) {
            // The reference's chain (annealing.py:121-146): every try mutates the state the previous
            // try left behind, is evaluated on its own (B = 1: the raster's 8-way split) and is
            // accepted or rejected before the next one is proposed.  Proposal number
            // (it - 1) * tries + t + 1 keys the random stream, as in the Python-driven loop.
            for (int t = 0; t < g->tries; ++t) {
                int rc = sa_batch(g, 1, sigma6, mutpb, log_scale_lo, log_scale_hi,
                                  (uint32_t)((size_t)(it - 1) * g->tries + t + 1), h_temperature[k],
                                  uniform + t, curve, st);
                if (rc) return rc;
            }
        } else {
            // `tries` independently mutated copies of the current state (annealing.py:121-128,
            // batched), one evaluation of all of them, the Metropolis tests applied in order
            int rc = sa_batch(g, g->tries, sigma6, mutpb, log_scale_lo, log_scale_hi, (uint32_t)it,
                              h_temperature[k], uniform, curve, st);
            if (rc) return rc;
        }
        g->iteration = it;
    }
    return sa_settle(g, st);  // the last batch's verdict: the state is final when this call returns
}

int ggs_sa_set_mode(ggs_sa *g, int batched_neighbours)
{
    if (!g) {
        set_error("ggs_sa_set_mode: NULL engine");
        return GGS_EINVAL;
    }
    g->sequential = (batched_neighbours == 0);
    return GGS_OK;
}

int ggs_sa_state(ggs_sa *g, void *stream, int *h_iteration, double *h_best_energy,
                 double *h_current_energy, double *h_curves2, int curves_from, float *h_best_state,
                 float *h_current_state)
{
    if (!g || g->iteration < 0) {
        set_error("ggs_sa_state: engine not started");
        return GGS_EINVAL;
    }
    if (curves_from < 0 || curves_from > g->iteration + 1) {
        set_error("ggs_sa_state: curves_from %d outside [0, %d]", curves_from, g->iteration + 1);
        return GGS_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(g->device);
    GGS_TRY(on_device.status());
    const size_t row = (size_t)g->N * 9 * sizeof(float);
    if (h_best_energy)
        GGS_TRY(cudaMemcpyAsync(h_best_energy, g->e_state + 2 * g->ep + 1, sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    if (h_current_energy)
        GGS_TRY(cudaMemcpyAsync(h_current_energy, g->e_state + 2 * g->ep, sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    if (h_curves2 && curves_from <= g->iteration)
        GGS_TRY(cudaMemcpyAsync(h_curves2, g->curves + (size_t)curves_from * 2,
                                (size_t)(g->iteration + 1 - curves_from) * 2 * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
    if (h_best_state) GGS_TRY(cudaMemcpyAsync(h_best_state, g->best, row, cudaMemcpyDeviceToHost, st));
    if (h_current_state)
        GGS_TRY(cudaMemcpyAsync(h_current_state, g->current, row, cudaMemcpyDeviceToHost, st));
    GGS_TRY(cudaStreamSynchronize(st));
    if (h_iteration) *h_iteration = g->iteration;
    return GGS_OK;
}

}  // extern "C"
