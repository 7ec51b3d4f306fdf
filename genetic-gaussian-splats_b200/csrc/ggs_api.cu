// extern "C" surface of libggs_b200.so (include/ggs_b200.h): argument checking, workspace
// carving, launch sequencing and the host-buffer path.  No kernels live here.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "ggs_common.cuh"

namespace ggs {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- run-time options: defaults from the environment (read once), changed by ggs_set_option ----
struct Options {
    int pdl;    // programmatic dependent launch between the kernels of a step (GGS_B200_PDL, default 1)
    int split;  // CTAs per (candidate, tile): 0 = automatic (GGS_B200_SPLIT)
    int tile_order;  // 1 (default): grids between two CTAs per SM and sixteen waves run the tiles centre-out
                     // (GGS_B200_TILE_ORDER); 0: candidate-major always.  Results do not depend on it.
    int fuse;   // decode fused into the raster: 0 = never (default), 1 = whenever a segment fits the
                // list, -1 = when the grid is a single wave and it fits (GGS_B200_FUSE).  Measured on
                // the B200 the fused variant never wins: every CTA pays the decode's chain of
                // transcendentals before its first blend, which the two-kernel path hides behind
                // the previous launch.
};
static int env_int(const char *name, int fallback)
{
    const char *v = getenv(name);
    return (v != nullptr && v[0] != '\0') ? atoi(v) : fallback;
}
static Options &options()
{
    static Options o = {env_int("GGS_B200_PDL", 1) != 0, env_int("GGS_B200_SPLIT", 0),
                        env_int("GGS_B200_TILE_ORDER", 1), env_int("GGS_B200_FUSE", 0)};
    return o;
}

bool pdl_enabled() { return options().pdl != 0; }

static int wave_slots()
{
    // CTAs of the raster resident at once on the current device: 8 per SM (64 registers x 128 threads)
    static int per_device[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148 * 8;
    if (per_device[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        per_device[dev] = sms * 8;
    }
    return per_device[dev];
}

int choose_split(int B, int N, int H, int W)
{
    const int forced = options().split;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
    // Measured on the B200 (tools/time_split_policy.py, profiles/r02_split_policy.txt): a cluster of 8
    // pays only while the grid stays within HALF a wave (one SA try at 256x256: 26.0 -> 13.8 us; two
    // candidates: 17.4 us at 8 against 15.9 at 4 -- every CTA folds `split` states); 2 and 4 pay up to
    // ~7 CTAs per SM (1,024 CTAs on 148 SMs) as long as the unsplit grid has less than 3 CTAs per SM
    // (3 candidates at 256x256: 20.9 us at 2, 18.1 at 4; 8 candidates = 512 CTAs: split 1 is the
    // fastest); segments keep at least 16 splats; and never on deep genomes, whose bands go opaque
    // early: the segments cannot see each other's saturation (512x512 / 4,000 splats, one frame:
    // 40.8 us at split 1, 68 us at split 4).
    if (N > 1536) return 1;
    const int64_t ctas = (int64_t)B * tiles_x(W) * tiles_y(H);
    const int slots = wave_slots();
    const auto segment_ok = [N](int k) { return (N + k - 1) / k >= 16; };
    if (ctas * 8 <= slots / 2 && segment_ok(8)) return 8;
    int k = 1;
    if (ctas * 8 < slots * 3)  // fewer than 3 CTAs per SM unsplit
        while (k < 4 && ctas * (k * 2) <= slots - slots / 8 && segment_ok(2 * k)) k *= 2;
    return k;
}

static int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return GGS_ECUDA;
}

#define GGS_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

// Workspace layout: [ticket counters | records | cull boxes | partials].  The counters come first
// so that their place does not depend on N, H, W: an owner that cleared the head of its buffer
// once can vouch for them whatever it evaluates next (EvalOptions::counters_zeroed).
size_t workspace_bytes(int B, int N, int H, int W)
{
    const size_t S = (size_t)B * (size_t)N;
    const size_t nt = (size_t)tiles_x(W) * tiles_y(H);
    size_t n = 0;
    n += align_up((size_t)B * sizeof(int), 256);
    n += align_up(S * sizeof(SplatRec), 256);
    n += align_up(S * sizeof(uint2), 256);
    n += align_up((size_t)B * nt * kMaxSplit * sizeof(float2), 256);
    return n;
}

Workspace carve_workspace(void *base, int B, int N, int H, int W)
{
    const size_t S = (size_t)B * (size_t)N;
    char *p = static_cast<char *>(base);
    Workspace ws;
    ws.counter = reinterpret_cast<int *>(p);
    p += align_up((size_t)B * sizeof(int), 256);
    ws.rec = reinterpret_cast<float4 *>(p);
    p += align_up(S * sizeof(SplatRec), 256);
    ws.aabb = reinterpret_cast<uint2 *>(p);
    p += align_up(S * sizeof(uint2), 256);
    ws.partial = reinterpret_cast<float2 *>(p);
    return ws;
}

static int check_shape(int B, int N, int cols, int H, int W)
{
    if (B < 0 || N < 0) {
        set_error("B and N must be non-negative (got B=%d N=%d)", B, N);
        return GGS_EINVAL;
    }
    if (cols < 9) {
        set_error("expected at least 9 genome cols (got %d)", cols);
        return GGS_EINVAL;
    }
    if (H < 1 || W < 1 || H > GGS_MAX_SIDE || W > GGS_MAX_SIDE) {
        set_error("H and W must be in [1, %d] (got H=%d W=%d)", GGS_MAX_SIDE, H, W);
        return GGS_EINVAL;
    }
    return GGS_OK;
}

static int check_layout(int layout)
{
    if (layout != GGS_LAYOUT_AXES_ANGLE && layout != GGS_LAYOUT_CHOLESKY) {
        set_error("unknown genome layout %d", layout);
        return GGS_EINVAL;
    }
    return GGS_OK;
}

// ---- optional per-kernel event log (ggs_timing_*) ------------------------------------
struct TimingLog {
    static constexpr int kCap = 4096;
    bool enabled = false;
    int used = 0;       // evaluations logged
    int created = 0;    // event triples created so far
    cudaEvent_t ev[kCap][3];
};
// Both belong to the calling thread: an evaluation issued by another thread (another device, an
// engine on its own stream) is neither timed nor switched to the instrumented kernel.
static thread_local TimingLog g_timing;
static thread_local unsigned long long *g_stats = nullptr;  // ggs_stats_target(): device counters or NULL
static thread_local int g_stats_device = -1;

static cudaEvent_t *timing_slot()
{
    if (!g_timing.enabled || g_timing.used >= TimingLog::kCap) return nullptr;
    if (g_timing.used == g_timing.created) {
        for (int k = 0; k < 3; ++k)
            if (cudaEventCreate(&g_timing.ev[g_timing.created][k]) != cudaSuccess) return nullptr;
        ++g_timing.created;
    }
    return g_timing.ev[g_timing.used++];
}

// decode + raster (or the fused raster alone) on `stream`; the one launch sequence behind every
// public entry.
int evaluate(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
             float k_sigma, const float bg[3], const float *d_target, const float *d_mask,
             int mode, float beta, float *d_fitness, void *d_images, int image_u8,
             void *d_workspace, size_t workspace_bytes_given, cudaStream_t stream,
             const EvalOptions &opt)
{
    if (B == 0) return GGS_OK;
    if (d_genomes == nullptr && N > 0) {
        set_error("d_genomes is NULL");
        return GGS_EINVAL;
    }
    const size_t need = workspace_bytes(B, N, H, W);
    if (d_workspace == nullptr || workspace_bytes_given < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes_given);
        return GGS_EWORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255u) != 0) {
        set_error("workspace must be 256-byte aligned");
        return GGS_EINVAL;
    }
    if (opt.split != 0 && opt.split != 1 && opt.split != 2 && opt.split != 4 && opt.split != 8) {
        set_error("split must be 0 (automatic), 1, 2, 4 or 8 (got %d)", opt.split);
        return GGS_EINVAL;
    }
    RasterLaunch q;
    q.ws = carve_workspace(d_workspace, B, N, H, W);
    q.B = B;
    q.N = N;
    q.H = H;
    q.W = W;
    q.bg[0] = bg[0];
    q.bg[1] = bg[1];
    q.bg[2] = bg[2];
    q.d_target = d_target;
    q.d_mask = d_mask;
    q.mode = mode;
    q.beta = beta;
    q.d_fitness = d_fitness;
    q.d_images = d_images;
    q.image_u8 = image_u8;
    q.d_genomes = d_genomes;
    q.layout = layout;
    q.cols = cols;
    q.k_sigma = k_sigma;
    q.peers = opt.peers;
    // the instrumented kernel exists for the throughput path only
    int dev = -1;
    const bool stats = g_stats != nullptr && cudaGetDevice(&dev) == cudaSuccess && dev == g_stats_device;
    q.d_stats = stats ? g_stats : nullptr;
    q.split = stats ? 1 : (opt.split != 0 ? opt.split : choose_split(B, N, H, W));
    const int fuse = stats ? 0 : (opt.fuse >= 0 ? opt.fuse : options().fuse);
    const int64_t ctas = (int64_t)B * tiles_x(W) * tiles_y(H) * q.split;
    q.small_grid = ctas <= wave_slots();
    // Centre-out CTA order (ggs_raster.cu, tile_geometry).  Measured: -4..-11 % on grids of one to
    // four waves (tools/time_cta_order.py), -8..-15 % on grids of LESS than a wave with more than one
    // CTA per SM -- the block scheduler deals CTAs to the SMs in index order, so with the heavy
    // tiles first the SMs that get one CTA more than the others get a light one
    // (tools/time_cta_order_small.py) -- still -4 % at 4.3 waves, -2.5 % at 7, -1 % at 14, nothing at
    // config 3's 55 (tools/time_cta_order_large.py); nothing on deep genomes, where saturation and
    // not the list length decides what a tile costs.
    q.interior_first = options().tile_order != 0 && N <= 1536 && ctas > wave_slots() / 8 &&
                       ctas <= 16 * (int64_t)wave_slots();
    q.fused = !opt.decoded && N > 0 && fuse != 0 && fused_decode_possible(N, q.split) &&
              (fuse == 1 || q.small_grid);

    cudaEvent_t *ev = timing_slot();
    if (ev) GGS_CUDA(cudaEventRecord(ev[0], stream));
    if (q.fused) {
        // no decode launch clears the ticket counters: do it here unless their owner vouches for them
        if (!opt.counters_zeroed) GGS_CUDA(cudaMemsetAsync(q.ws.counter, 0, (size_t)B * sizeof(int), stream));
    } else if (!opt.decoded) {
        GGS_CUDA(launch_decode(d_genomes, layout, (int64_t)B * N, cols, H, W, k_sigma, q.ws.rec, q.ws.aabb,
                               nullptr, nullptr, q.ws.counter, B, stream));
    }
    if (ev) GGS_CUDA(cudaEventRecord(ev[1], stream));
    GGS_CUDA(launch_raster(q, stream));
    if (ev) GGS_CUDA(cudaEventRecord(ev[2], stream));
    return GGS_OK;
}

}  // namespace ggs

using namespace ggs;

extern "C" {

int ggs_abi_version(void) { return GGS_ABI_VERSION; }

const char *ggs_last_error(void) { return g_err; }

int ggs_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

size_t ggs_workspace_bytes(int B, int N, int H, int W)
{
    if (B < 0 || N < 0 || H < 1 || W < 1) return 0;
    return workspace_bytes(B, N, H, W);
}

int ggs_encode(const float *d_axes, int64_t rows, int cols, float *d_chol, void *stream)
{
    if (rows < 0 || cols < 9 || (rows > 0 && (d_axes == nullptr || d_chol == nullptr))) {
        set_error("ggs_encode: bad arguments (rows=%lld cols=%d)", (long long)rows, cols);
        return GGS_EINVAL;
    }
    GGS_CUDA(launch_encode(d_axes, rows, cols, d_chol, static_cast<cudaStream_t>(stream)));
    return GGS_OK;
}

int ggs_decode(const float *d_genomes, int layout, int64_t rows, int cols, int H, int W,
               float k_sigma, float *d_out_f, int32_t *d_out_i, void *stream)
{
    int rc = check_layout(layout);
    if (rc) return rc;
    rc = check_shape(0, 0, cols, H, W);
    if (rc) return rc;
    if (rows < 0 || (rows > 0 && (d_genomes == nullptr || d_out_f == nullptr || d_out_i == nullptr))) {
        set_error("ggs_decode: bad arguments");
        return GGS_EINVAL;
    }
    GGS_CUDA(launch_decode(d_genomes, layout, rows, cols, H, W, k_sigma, nullptr, nullptr, d_out_f,
                           d_out_i, nullptr, 0, static_cast<cudaStream_t>(stream)));
    return GGS_OK;
}

int ggs_render(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
               float k_sigma, const float *h_background, float *d_images, void *d_workspace,
               size_t workspace_bytes_given, void *stream)
{
    int rc = check_layout(layout);
    if (rc) return rc;
    rc = check_shape(B, N, cols, H, W);
    if (rc) return rc;
    if (B > 0 && d_images == nullptr) {
        set_error("ggs_render: d_images is NULL");
        return GGS_EINVAL;
    }
    const float white[3] = {1.0f, 1.0f, 1.0f};
    const float *bg = h_background ? h_background : white;
    return evaluate(d_genomes, layout, B, N, cols, H, W, k_sigma, bg, nullptr, nullptr,
                    GGS_MODE_PLAIN, 1.0f, nullptr, d_images, 0, d_workspace, workspace_bytes_given,
                    static_cast<cudaStream_t>(stream));
}

int ggs_render_u8(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
                  float k_sigma, const float *h_background, unsigned char *d_images_u8,
                  void *d_workspace, size_t workspace_bytes_given, void *stream)
{
    int rc = check_layout(layout);
    if (rc) return rc;
    rc = check_shape(B, N, cols, H, W);
    if (rc) return rc;
    if (B > 0 && d_images_u8 == nullptr) {
        set_error("ggs_render_u8: d_images_u8 is NULL");
        return GGS_EINVAL;
    }
    const float white[3] = {1.0f, 1.0f, 1.0f};
    const float *bg = h_background ? h_background : white;
    return evaluate(d_genomes, layout, B, N, cols, H, W, k_sigma, bg, nullptr, nullptr,
                    GGS_MODE_PLAIN, 1.0f, nullptr, d_images_u8, 1, d_workspace,
                    workspace_bytes_given, static_cast<cudaStream_t>(stream));
}

int ggs_fitness_ex(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
                   float k_sigma, const float *d_target, const float *d_mask, int mode,
                   float boost_beta, float *d_fitness, float *d_images, void *d_workspace,
                   size_t workspace_bytes_given, int split, void *stream)
{
    int rc = check_layout(layout);
    if (rc) return rc;
    rc = check_shape(B, N, cols, H, W);
    if (rc) return rc;
    if (mode != GGS_MODE_PLAIN && mode != GGS_MODE_MASK && mode != GGS_MODE_BOOST) {
        set_error("unknown fitness mode %d", mode);
        return GGS_EINVAL;
    }
    if (B > 0 && (d_target == nullptr || d_fitness == nullptr)) {
        set_error("ggs_fitness: d_target / d_fitness is NULL");
        return GGS_EINVAL;
    }
    if (mode != GGS_MODE_PLAIN && d_mask == nullptr) {
        set_error("ggs_fitness: mode %d needs a weight mask", mode);
        return GGS_EINVAL;
    }
    const float white[3] = {1.0f, 1.0f, 1.0f};  // render.py:209
    EvalOptions opt;
    opt.split = split;
    return evaluate(d_genomes, layout, B, N, cols, H, W, k_sigma, white, d_target, d_mask, mode,
                    boost_beta, d_fitness, d_images, 0, d_workspace, workspace_bytes_given,
                    static_cast<cudaStream_t>(stream), opt);
}

int ggs_fitness(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
                float k_sigma, const float *d_target, const float *d_mask, int mode,
                float boost_beta, float *d_fitness, float *d_images, void *d_workspace,
                size_t workspace_bytes_given, void *stream)
{
    return ggs_fitness_ex(d_genomes, layout, B, N, cols, H, W, k_sigma, d_target, d_mask, mode,
                          boost_beta, d_fitness, d_images, d_workspace, workspace_bytes_given, 0, stream);
}

/* ---------------------------------------------------------------------------------- */
/* host-buffer path                                                                    */

constexpr int kMaxSlices = 8;

struct ggs_ctx {
    int device = 0;
    cudaStream_t copy = nullptr;                // host -> device genome slices, running ahead
    cudaStream_t stream[2] = {nullptr, nullptr};  // compute, alternating per slice
    cudaEvent_t landed[kMaxSlices] = {};        // slice k is on the device
    cudaEvent_t done[2] = {nullptr, nullptr};
    float *d_target = nullptr;
    float *d_mask = nullptr;
    int H = 0, W = 0;
    bool has_mask = false;
    // grow-only device buffers: the whole population's genomes, one workspace per slice
    float *d_genomes = nullptr;
    size_t genomes_cap = 0;
    void *d_ws[kMaxSlices] = {};
    size_t ws_cap[kMaxSlices] = {};
    float *d_fitness = nullptr;
    size_t fitness_cap = 0;
};

static int grow(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return GGS_OK;
    if (*p) GGS_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    GGS_CUDA(cudaMalloc(p, need));
    *cap = need;
    return GGS_OK;
}

int ggs_ctx_create(int device, ggs_ctx **out)
{
    if (out == nullptr) {
        set_error("ggs_ctx_create: out is NULL");
        return GGS_EINVAL;
    }
    *out = nullptr;
    int n = 0;
    GGS_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error("ggs_ctx_create: device %d not visible (%d devices)", device, n);
        return GGS_ENODEVICE;
    }
    DeviceGuard on_device(device);
    GGS_CUDA(on_device.status());
    ggs_ctx *c = new (std::nothrow) ggs_ctx();
    if (!c) {
        set_error("out of host memory");
        return GGS_EINVAL;
    }
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < kMaxSlices && e == cudaSuccess; ++i)
        e = cudaEventCreateWithFlags(&c->landed[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        ggs_ctx_destroy(c);  // releases whatever was created
        return cuda_fail(e, "ggs_ctx_create: stream / event creation");
    }
    *out = c;
    return GGS_OK;
}

void ggs_ctx_destroy(ggs_ctx *c)
{
    if (!c) return;
    DeviceGuard on_device(c->device);
    if (c->copy) cudaStreamSynchronize(c->copy);
    for (int i = 0; i < 2; ++i) {
        if (c->stream[i]) cudaStreamSynchronize(c->stream[i]);
        if (c->done[i]) cudaEventDestroy(c->done[i]);
        if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
    }
    for (int i = 0; i < kMaxSlices; ++i) {
        if (c->landed[i]) cudaEventDestroy(c->landed[i]);
        if (c->d_ws[i]) cudaFree(c->d_ws[i]);
    }
    if (c->copy) cudaStreamDestroy(c->copy);
    if (c->d_genomes) cudaFree(c->d_genomes);
    if (c->d_target) cudaFree(c->d_target);
    if (c->d_mask) cudaFree(c->d_mask);
    if (c->d_fitness) cudaFree(c->d_fitness);
    delete c;
}

int ggs_ctx_set_target(ggs_ctx *c, const float *h_target, const float *h_mask, int H, int W)
{
    if (!c || !h_target) {
        set_error("ggs_ctx_set_target: NULL argument");
        return GGS_EINVAL;
    }
    int rc = check_shape(0, 0, 9, H, W);
    if (rc) return rc;
    DeviceGuard on_device(c->device);
    GGS_CUDA(on_device.status());
    GGS_CUDA(cudaStreamSynchronize(c->copy));
    for (int i = 0; i < 2; ++i) GGS_CUDA(cudaStreamSynchronize(c->stream[i]));
    if (c->d_target) GGS_CUDA(cudaFree(c->d_target));
    if (c->d_mask) GGS_CUDA(cudaFree(c->d_mask));
    c->d_target = c->d_mask = nullptr;
    const size_t P = (size_t)H * W;
    GGS_CUDA(cudaMalloc(&c->d_target, P * 3 * sizeof(float)));
    GGS_CUDA(cudaMemcpy(c->d_target, h_target, P * 3 * sizeof(float), cudaMemcpyHostToDevice));
    c->has_mask = (h_mask != nullptr);
    if (h_mask) {
        GGS_CUDA(cudaMalloc(&c->d_mask, P * sizeof(float)));
        GGS_CUDA(cudaMemcpy(c->d_mask, h_mask, P * sizeof(float), cudaMemcpyHostToDevice));
    }
    c->H = H;
    c->W = W;
    return GGS_OK;
}

int ggs_ctx_fitness_host(ggs_ctx *c, const float *h_genomes, int layout, int B, int N, int cols,
                         float k_sigma, int mode, float boost_beta, float *h_fitness)
{
    if (!c || c->d_target == nullptr) {
        set_error("ggs_ctx_fitness_host: call ggs_ctx_set_target first");
        return GGS_EINVAL;
    }
    int rc = check_layout(layout);
    if (rc) return rc;
    rc = check_shape(B, N, cols, c->H, c->W);
    if (rc) return rc;
    if (mode != GGS_MODE_PLAIN && !c->has_mask) {
        set_error("ggs_ctx_fitness_host: mode %d needs a mask in ggs_ctx_set_target", mode);
        return GGS_EINVAL;
    }
    if (B == 0) return GGS_OK;
    if (!h_genomes || !h_fitness) {
        set_error("ggs_ctx_fitness_host: NULL buffer");
        return GGS_EINVAL;
    }
    DeviceGuard on_device(c->device);
    GGS_CUDA(on_device.status());

    // The genomes go up in slices on a copy stream that runs ahead of the compute streams;
    // slice k is evaluated as soon as it has landed.  Sizes double (B/64, B/32, B/16, B/8, rest),
    // so only a sixty-fourth of the copy is exposed and every later copy hides behind the previous,
    // smaller, evaluation; consecutive slices use alternating compute streams so the tail of
    // one raster overlaps the head of the next.  (Measured at config 3: first slice B/16 with
    // four slices 3.200 ms per call, B/32 3.173, B/64 with five slices 3.166, B/128 3.175.)
    int start[kMaxSlices + 1];
    int ns = 0;
    start[0] = 0;
    for (int sz = std::max(16, B / 64); start[ns] < B && ns < kMaxSlices; sz *= 2) {
        const bool last = (ns == 4) || (ns == kMaxSlices - 1) || start[ns] + sz >= B;
        start[ns + 1] = last ? B : start[ns] + sz;
        ++ns;
    }
    const size_t row_bytes = (size_t)N * cols * sizeof(float);
    rc = grow(reinterpret_cast<void **>(&c->d_fitness), &c->fitness_cap, (size_t)B * sizeof(float));
    if (rc) return rc;
    rc = grow(reinterpret_cast<void **>(&c->d_genomes), &c->genomes_cap,
              std::max<size_t>((size_t)B * row_bytes, 256));
    if (rc) return rc;
    for (int k = 0; k < ns; ++k) {
        const size_t had = c->ws_cap[k];
        rc = grow(&c->d_ws[k], &c->ws_cap[k], workspace_bytes(start[k + 1] - start[k], N, c->H, c->W));
        if (rc) return rc;
        // a fresh buffer: clear it once, every evaluation leaves its ticket counters at zero
        if (c->ws_cap[k] != had) GGS_CUDA(cudaMemsetAsync(c->d_ws[k], 0, c->ws_cap[k], c->stream[k & 1]));
    }
    const float white[3] = {1.0f, 1.0f, 1.0f};
    // One kernel configuration for the whole call, chosen from the WHOLE batch: the result does
    // not depend on how the batch is sliced and equals ggs_fitness on the same genomes.
    EvalOptions opt;
    opt.split = choose_split(B, N, c->H, c->W);
    // evaluate()'s own rule (option `fuse`: 0 = never, the default; 1 = whenever a segment fits;
    // -1 = on single-wave grids), applied to the whole batch instead of slice by slice
    const int fuse_option = options().fuse;
    opt.fuse = (fuse_option != 0 && fused_decode_possible(N, opt.split) &&
                (fuse_option == 1 ||
                 (int64_t)B * tiles_x(c->W) * tiles_y(c->H) * opt.split <= wave_slots())) ? 1 : 0;
    opt.counters_zeroed = true;
    for (int k = 0; k < ns; ++k) {
        const size_t off = (size_t)start[k] * N * cols;
        GGS_CUDA(cudaMemcpyAsync(c->d_genomes + off, h_genomes + off,
                                 (size_t)(start[k + 1] - start[k]) * row_bytes,
                                 cudaMemcpyHostToDevice, c->copy));
        GGS_CUDA(cudaEventRecord(c->landed[k], c->copy));
    }
    for (int k = 0; k < ns; ++k) {
        const int s = k & 1, sb = start[k + 1] - start[k];
        GGS_CUDA(cudaStreamWaitEvent(c->stream[s], c->landed[k], 0));
        rc = evaluate(c->d_genomes + (size_t)start[k] * N * cols, layout, sb, N, cols, c->H, c->W,
                      k_sigma, white, c->d_target, c->d_mask, mode, boost_beta,
                      c->d_fitness + start[k], nullptr, 0, c->d_ws[k], c->ws_cap[k], c->stream[s], opt);
        if (rc) return rc;
    }
    // Drain: stream 1's work must finish before the single D2H issued on stream 0.
    GGS_CUDA(cudaEventRecord(c->done[1], c->stream[1]));
    GGS_CUDA(cudaStreamWaitEvent(c->stream[0], c->done[1], 0));
    GGS_CUDA(cudaMemcpyAsync(h_fitness, c->d_fitness, (size_t)B * sizeof(float),
                             cudaMemcpyDeviceToHost, c->stream[0]));
    GGS_CUDA(cudaStreamSynchronize(c->stream[0]));
    return GGS_OK;
}

int ggs_ga_breed(const float *d_population, const float *d_fitness, int P, int N, int cols,
                 float *d_offspring, int tour_k, float cxpb, float mutpb, const float *h_sigma6,
                 float log_scale_lo, float log_scale_hi, uint64_t seed, uint32_t generation,
                 void *stream)
{
    if (P < 0 || N < 0 || cols < 9 || tour_k < 1 || h_sigma6 == nullptr) {
        set_error("ggs_ga_breed: bad arguments (P=%d N=%d cols=%d tour_k=%d)", P, N, cols, tour_k);
        return GGS_EINVAL;
    }
    if (P > 0 && N > 0 && (!d_population || !d_fitness || !d_offspring)) {
        set_error("ggs_ga_breed: NULL buffer");
        return GGS_EINVAL;
    }
    if (d_population == d_offspring && P > 0) {
        set_error("ggs_ga_breed: offspring must not alias the population");
        return GGS_EINVAL;
    }
    GGS_CUDA(launch_breed(d_population, d_fitness, P, N, cols, P, d_offspring, tour_k, cxpb, mutpb,
                          h_sigma6, log_scale_lo, log_scale_hi, seed, generation,
                          static_cast<cudaStream_t>(stream)));
    return GGS_OK;
}

size_t ggs_mask_workspace_bytes(int H, int W)
{
    if (H <= 0 || W <= 0) return 0;
    return ggs::mask_workspace_bytes(H, W);
}

int ggs_importance_mask(const float *d_image, int H0, int W0, int H, int W, int image_is_0_255,
                        const int *h_edge_scales, int n_scales, double w_edge, double w_var,
                        double gamma, double floor, int smooth, double strength, float *d_mask,
                        void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (H0 <= 0 || W0 <= 0 || H <= 0 || W <= 0 || H0 > GGS_MAX_SIDE || W0 > GGS_MAX_SIDE ||
        H > GGS_MAX_SIDE || W > GGS_MAX_SIDE || n_scales < 0 || (n_scales > 0 && !h_edge_scales) ||
        smooth < 0 || (smooth > 0 && (smooth & 1) == 0)) {
        set_error("ggs_importance_mask: bad arguments (source %dx%d, work %dx%d, %d scales, "
                  "smooth %d: must be 0 or odd)", H0, W0, H, W, n_scales, smooth);
        return GGS_EINVAL;
    }
    for (int k = 0; k < n_scales; ++k) {
        const int s = h_edge_scales[k];
        if (s < 1 || s > H || s > W) {  // avg_pool2d would produce an empty map (mask.py:54)
            set_error("ggs_importance_mask: edge scale %d does not fit a %dx%d image", s, H, W);
            return GGS_EINVAL;
        }
    }
    if (!d_image || !d_mask || !d_workspace) {
        set_error("ggs_importance_mask: NULL buffer");
        return GGS_EINVAL;
    }
    if (workspace_bytes < ggs::mask_workspace_bytes(H, W)) {
        set_error("ggs_importance_mask: workspace of %zu bytes, %zu needed", workspace_bytes,
                  ggs::mask_workspace_bytes(H, W));
        return GGS_EWORKSPACE;
    }
    // Scalar arithmetic in double, then one rounding to float32: what Python + torch do.
    GGS_CUDA(ggs::launch_importance_mask(
        d_image, H0, W0, H, W, image_is_0_255 != 0, h_edge_scales, n_scales, (float)w_edge,
        (float)w_var, (float)gamma, (float)(1.0 - floor), (float)floor, smooth,
        (float)(1.0 - strength), (float)strength, strength < 1.0 ? 1 : 0, d_mask, d_workspace,
        static_cast<cudaStream_t>(stream)));
    return GGS_OK;
}

int ggs_stats_target(unsigned long long *d_counters2)
{
    g_stats = d_counters2;
    g_stats_device = -1;
    if (d_counters2 != nullptr) GGS_CUDA(cudaGetDevice(&g_stats_device));
    return GGS_OK;
}

int ggs_set_option(const char *name, int value)
{
    if (name != nullptr && strcmp(name, "pdl") == 0) {
        options().pdl = value != 0;
    } else if (name != nullptr && strcmp(name, "split") == 0 &&
               (value == 0 || value == 1 || value == 2 || value == 4 || value == 8)) {
        options().split = value;
    } else if (name != nullptr && strcmp(name, "fuse") == 0 && value >= -1 && value <= 1) {
        options().fuse = value;
    } else if (name != nullptr && strcmp(name, "tile_order") == 0 && (value == 0 || value == 1)) {
        options().tile_order = value;
    } else {
        set_error("ggs_set_option: unknown option or value (%s = %d)", name ? name : "(null)", value);
        return GGS_EINVAL;
    }
    return GGS_OK;
}

int ggs_tile_order(int ntx, int nty, int *out_xy)
{
    if (ntx < 1 || nty < 1 || out_xy == nullptr) {
        set_error("ggs_tile_order: bad arguments (ntx=%d nty=%d)", ntx, nty);
        return GGS_EINVAL;
    }
    for (int r = 0; r < ntx * nty; ++r) centre_out_tile(r, ntx, nty, out_xy[2 * r], out_xy[2 * r + 1]);
    return GGS_OK;
}

int ggs_choose_split(int B, int N, int H, int W)
{
    if (B < 1 || N < 0 || H < 1 || W < 1) return 1;
    return choose_split(B, N, H, W);
}

int ggs_timing_enable(int enable)
{
    g_timing.enabled = (enable != 0);
    g_timing.used = 0;
    return GGS_OK;
}

int ggs_timing_read(float *h_decode_ms, float *h_raster_ms, int *h_evaluations)
{
    double dec = 0.0, ras = 0.0;
    for (int i = 0; i < g_timing.used; ++i) {
        GGS_CUDA(cudaEventSynchronize(g_timing.ev[i][2]));
        float a = 0.0f, b = 0.0f;
        GGS_CUDA(cudaEventElapsedTime(&a, g_timing.ev[i][0], g_timing.ev[i][1]));
        GGS_CUDA(cudaEventElapsedTime(&b, g_timing.ev[i][1], g_timing.ev[i][2]));
        dec += a;
        ras += b;
    }
    if (h_decode_ms) *h_decode_ms = (float)dec;
    if (h_raster_ms) *h_raster_ms = (float)ras;
    if (h_evaluations) *h_evaluations = g_timing.used;
    g_timing.used = 0;
    return GGS_OK;
}

int ggs_probe_peaks(float *h_out5)
{
    if (!h_out5) {
        set_error("ggs_probe_peaks: NULL output");
        return GGS_EINVAL;
    }
    GGS_CUDA(probe_peaks(h_out5));
    return GGS_OK;
}

}  // extern "C"
