// Shared definitions of the sm_100a render + fitness path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ggs_b200.h"

namespace ggs {

// Raster geometry: one CTA per (candidate, 32x32 tile); one warp per 32x8 band, lane = pixel
// column, each thread owns 8 vertically adjacent pixels.  With this mapping the AABB row test
// is warp-uniform (no per-pixel predicate) and the AABB column test is one select per
// (thread, splat) folded into the exponent.
constexpr int kTileW = 32;
#ifndef GGS_ROWS
#define GGS_ROWS 8    // pixel rows per thread: 8 or 16 (row codes hold 4-bit row indices)
#endif
#ifndef GGS_WARPS
#define GGS_WARPS 4   // warps (bands) per CTA, <= 4 (one row-code byte per band)
#endif
static_assert(GGS_ROWS == 8 || GGS_ROWS == 16, "rows per thread must be 8 or 16");
static_assert(GGS_WARPS >= 1 && GGS_WARPS <= 4, "1..4 warps per CTA");
constexpr int kRowsPerThread = GGS_ROWS;
constexpr int kWarps = GGS_WARPS;
constexpr int kTileH = kWarps * kRowsPerThread;
constexpr int kThreads = kWarps * 32;
#ifndef GGS_LIST_CAP
#define GGS_LIST_CAP 512
#endif
constexpr int kListCap = GGS_LIST_CAP;  // staged splat records per flush (48 B each)

// Latency path: up to kMaxSplit CTAs (one thread-block cluster) share a (candidate, tile), each
// compositing one segment of the genome (ggs_raster.cu, raster_split_kernel).
constexpr int kMaxSplit = 8;   // portable cluster size limit
static_assert(kTileH % kMaxSplit == 0, "every CTA of a cluster finishes an equal share of the tile's rows");

constexpr int kDecodeThreads = 256;
constexpr int kDecodeStageMaxCols = 16;

// Decoded splat record, 48 B = 3 x float4, stored [B][N] in the workspace.
//   e(X,Y) = A*qx^2 + Bq*qx*qy + Cq*qy^2 + la,  f = 2^e  ( = exp(-quad/2) * alpha )
// with A = -0.5*log2(e)*sxx, Bq = -log2(e)*sxy, Cq = -0.5*log2(e)*syy, la = log2(alpha).
struct __align__(16) SplatRec {
    float cx, cy, A, Bq;
    float Cq, la, r, g;
    float b;
    int xpack;  // x0 | x1 << 16   (inclusive AABB, render.py:27-28)
    int ypack;  // y0 | y1 << 16   (render.py:29-30)
    float h;    // 2^(8*Cq) for the column recurrence, or -1: steep splat, exact path only
};
static_assert(sizeof(SplatRec) == 48, "SplatRec must be 3 float4");

struct Workspace {
    float4 *rec;      // [B*N*3]
    uint2 *aabb;      // [B*N]   x0|x1<<16, y0|y1<<16 (int16 each)
    float2 *partial;  // [B*ntiles*kMaxSplit] (numerator, denominator) per CTA
    int *counter;     // [B] CTAs finished per candidate (zero between launches)
};

// Fitness stores into the gathered vectors of the other GPUs of the box (ggs_peers.cu): the CTA
// that finishes a candidate writes its fitness straight into every rank's peer-mapped buffer, and
// the CTA that finishes the launch's last candidate raises this rank's flag on every rank.  No
// collective launch follows the raster.  n == 0 switches it off.
constexpr int kMaxPeers = 8;
struct PeerStores {
    int n = 0;                    // ranks that receive the values (world size), 0 = off
    int rank = 0;                 // this rank
    int offset = 0;               // index of this launch's candidate 0 in the gathered vector
    unsigned epoch = 0;           // value of the flags once the whole launch has been stored
    float *fit[kMaxPeers] = {};      // rank r's gathered vector of this epoch (own buffer for r == rank)
    unsigned *flag[kMaxPeers] = {};  // rank r's arrival flags, one per sender
    int *done = nullptr;          // candidates published by this launch (local, self-resetting)
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Tile of rank r in centre-out order (the CTA order of grids between two CTAs per SM and four
// waves, ggs_raster.cu): ring after ring from the innermost, and inside a ring the edge cells
// ahead of its four corners.  Ring k = the cells at distance k from the image border; a tile nearer
// the border is overlapped by fewer splats.  A permutation of the ntx x nty tiles for every size.
__host__ __device__ inline void centre_out_tile(int r, int ntx, int nty, int &tx, int &ty)
{
    int k = ((ntx < nty ? ntx : nty) - 1) >> 1;  // innermost ring
    int inner = 0;                               // cells strictly inside ring k
    for (; k > 0; --k) {
        const int here = (ntx - 2 * k) * (nty - 2 * k);
        if (r < here) break;
        inner = here;
    }
    const int w = ntx - 2 * k, h = nty - 2 * k;
    int j = r - inner;  // position inside ring k (a w x h frame at offset (k, k))
    if (w == 1 || h == 1) {  // a line, not a frame
        tx = k + (h == 1 ? j : 0);
        ty = k + (h == 1 ? 0 : j);
    } else if (j < 2 * (w - 2)) {  // top and bottom edges without their corners
        const int top = j < w - 2;
        tx = k + 1 + (top ? j : j - (w - 2));
        ty = top ? k : k + h - 1;
    } else if ((j -= 2 * (w - 2)) < 2 * (h - 2)) {  // left and right edges without their corners
        tx = (j & 1) ? k + w - 1 : k;
        ty = k + 1 + (j >> 1);
    } else {  // the four corners
        j -= 2 * (h - 2);
        tx = (j & 1) ? k + w - 1 : k;
        ty = (j & 2) ? k + h - 1 : k;
    }
}

inline int tiles_x(int W) { return (W + kTileW - 1) / kTileW; }
inline int tiles_y(int H) { return (H + kTileH - 1) / kTileH; }

size_t workspace_bytes(int B, int N, int H, int W);
Workspace carve_workspace(void *base, int B, int N, int H, int W);

// decode.cu
cudaError_t launch_decode(const float *d_genomes, int layout, int64_t rows, int cols, int H, int W,
                          float k_sigma, float4 *rec, uint2 *aabb, float *raw_f, int32_t *raw_i,
                          int *counters, int n_counters, cudaStream_t stream);
cudaError_t launch_encode(const float *d_axes, int64_t rows, int cols, float *d_chol,
                          cudaStream_t stream);

// raster.cu
struct RasterLaunch {
    Workspace ws;
    int B = 0, N = 0, H = 0, W = 0;
    float bg[3] = {1.0f, 1.0f, 1.0f};
    const float *d_target = nullptr, *d_mask = nullptr;
    int mode = GGS_MODE_PLAIN;
    float beta = 1.0f;
    float *d_fitness = nullptr;
    void *d_images = nullptr;
    int image_u8 = 0;
    unsigned long long *d_stats = nullptr;
    int split = 1;        // CTAs per (candidate, tile): 1, 2, 4 or 8
    bool fused = false;   // decode inside the raster (needs the genomes, ceil(N / split) <= kListCap)
    bool small_grid = false;  // at most one wave of CTAs: latency matters more than L1 traffic
    bool interior_first = false;  // two CTAs per SM to a few waves: tile-major, tiles from the centre outwards
    const float *d_genomes = nullptr;
    int layout = GGS_LAYOUT_AXES_ANGLE, cols = 9;
    float k_sigma = 3.0f;
    PeerStores peers;
};
cudaError_t launch_raster(const RasterLaunch &q, cudaStream_t stream);
bool fused_decode_possible(int N, int split);

// breed.cu
// Children [0, n_children) of the step are produced (counter-based streams: child c is the same
// whatever n_children is); a GA generation defines P of them, SA asks for `tries` mutated
// copies of a one-individual population.
cudaError_t launch_breed(const float *d_pop, const float *d_fitness, int P, int N, int cols,
                         int n_children, float *d_offspring, int tour_k, float cxpb, float mutpb,
                         const float sigma6[6], float log_lo, float log_hi, uint64_t seed,
                         uint32_t generation, cudaStream_t stream);

// Simulated annealing's proposal step: n_children mutated copies of one parent (no selection, no
// crossover: the same bits launch_breed gives for a one-individual population) AND their decoded
// records, written into the workspace of the evaluation that follows (EvalOptions::decoded).
//
// The kernel also JUDGES the evaluation that preceded it (annealing.py:129-146), so that a try is
// two launches -- propose, raster -- and not four: with judge.tries > 0 every CTA replays the
// Metropolis tests on the `tries` energies of the previous candidates (accept when dE <= 0 or
// u < exp(-dE / T); the best-so-far test follows every try) and mutates the state they leave --
// an accepted candidate of the previous batch, or the current state -- while CTA 0 records the
// outcome: energies (read from e_in, written to e_out: the other CTAs still read e_in), the curve
// point, the current / best rows.  n_children = 0 only judges (the trailing launch of a block).
constexpr int kMaxTries = 64;
struct ProposeJudge {
    int tries = 0;                    // 0: nothing to judge
    const float *energy = nullptr;    // [tries] fitness of the previous candidates
    const float *cand_prev = nullptr; // [tries][N][9] the previous candidates
    float *current = nullptr;         // [N][9] in / out
    float *best = nullptr;            // [N][9] in / out
    const double *e_in = nullptr;     // {e_current, e_best} before
    double *e_out = nullptr;          // {e_current, e_best} after
    double *curve = nullptr;          // (best, current) of the judged iteration
    double temperature = 0.0;
    double uniform[kMaxTries] = {};   // one U[0,1) draw per try (used when the try is uphill)
};
bool propose_possible(int N, int cols);
cudaError_t launch_propose(const float *d_parent, int N, int cols, int n_children, float *d_children,
                           float mutpb, const float sigma6[6], float log_lo, float log_hi,
                           uint64_t seed, uint32_t generation, const Workspace &ws, int H, int W,
                           float k_sigma, const ProposeJudge &judge, cudaStream_t stream);

size_t mask_workspace_bytes(int H, int W);
cudaError_t launch_importance_mask(const float *d_image, int H0, int W0, int H, int W, int div255,
                                   const int *scales, int n_scales, float w_edge, float w_var,
                                   float gamma, float one_minus_floor, float floor_,
                                   int smooth, float one_minus_strength, float strength,
                                   int blend, float *d_mask, void *d_ws, cudaStream_t st);

// probe.cu
cudaError_t probe_peaks(float *h_out5);

// peers.cu: the P2P fitness exchange between the GPUs of a box (struct ggs_peers is the C ABI's)
}  // namespace ggs
struct ggs_peers;
namespace ggs {
int peers_rank(const ggs_peers *p);
int peers_world(const ggs_peers *p);
int peers_capacity(const ggs_peers *p);
bool peers_ready(const ggs_peers *p);
PeerStores peers_next(ggs_peers *p, int offset);   // takes the next epoch number
unsigned *peers_flags(ggs_peers *p);               // this rank's arrival flags [world]
int *peers_status(ggs_peers *p);
float *peers_gathered(ggs_peers *p, unsigned epoch);  // this rank's gathered vector of that epoch
cudaError_t peers_signal_empty(const PeerStores &s, cudaStream_t st);
cudaError_t peers_wait(ggs_peers *p, unsigned epoch, cudaStream_t st);

void set_error(const char *fmt, ...);

// Entry points that own a device (contexts, engines, peers) run on it and hand the caller's
// current device back when they return.
class DeviceGuard {
public:
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
        status_ = (prev_ == device) ? cudaSuccess : cudaSetDevice(device);
        changed_ = (status_ == cudaSuccess && prev_ != device);
    }
    ~DeviceGuard()
    {
        if (changed_ && prev_ >= 0) cudaSetDevice(prev_);
    }
    cudaError_t status() const { return status_; }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;

private:
    int prev_ = -1;
    bool changed_ = false;
    cudaError_t status_ = cudaSuccess;
};

// ---- programmatic dependent launch (PDL) --------------------------------------------------
// The kernels of an evaluation (decode -> raster) and of a GA / SA step (breed -> decode ->
// raster -> select) are short at small populations, so the few microseconds between dependent
// launches count.  Every kernel launched through launch_kernel() may be scheduled while its
// predecessor in the stream is still draining; it executes pdl_wait() -- which returns once the
// predecessor has completed and its writes are visible -- before it touches global memory, and
// pdl_trigger() right after, so the same holds for its own successor.  Kernels that precede it
// and know nothing of this (torch's) simply trigger at exit.  GGS_B200_PDL=0 switches back to
// plain stream order (for A/B timing).
bool pdl_enabled();  // api.cu

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// A fitness value has been computed for candidate `b` of a launch of B candidates.
__device__ __forceinline__ void peer_publish(const PeerStores &p, int b, float fit, int B)
{
    if (p.n == 0) return;
    for (int r = 0; r < p.n; ++r)
        asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.fit[r] + p.offset + b), "f"(fit) : "memory");
    __threadfence_system();
    if (atomicAdd(p.done, 1) == B - 1) {  // every candidate of this launch is stored everywhere
        atomicExch(p.done, 0);
        __threadfence_system();
        for (int r = 0; r < p.n; ++r)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag[r] + p.rank), "r"(p.epoch) : "memory");
    }
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Thread r < world waits until sender r's arrival flag has reached `epoch` (the values that
// sender published before it are then visible).  A sender that never arrives -- a rank died --
// would spin for ever: give up after 20 s and leave a mark the host reads (ggs_peers_status).
__device__ __forceinline__ void peer_wait(const unsigned *flags, int world, unsigned epoch, int *status)
{
    const int r = threadIdx.x;
    if (r < world) {
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(flags + r) - epoch) < 0) {
            if (global_ns() - t0 > 20ull * 1000 * 1000 * 1000) {
                *status = 1;
                break;
            }
            __nanosleep(64);
        }
    }
}

// Same, for a launch that runs as thread-block clusters of `cluster` CTAs along x.
template <typename... P, typename... A>
inline cudaError_t launch_kernel_cluster(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem,
                                         int cluster, cudaStream_t stream, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

template <typename... P, typename... A>
inline cudaError_t launch_kernel(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, A... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
#endif

// api.cu: the launch sequence behind every evaluation entry -- decode + raster, or the raster's
// fused-decode variant alone -- on `stream`.
struct EvalOptions {
    int split = 0;   // CTAs per (candidate, tile): 0 = choose_split() for this B, else 1 / 2 / 4 / 8
    int fuse = -1;   // decode inside the raster: -1 = when it pays (one wave) and fits, 0 = never, 1 = if it fits
    bool counters_zeroed = false;  // the workspace's ticket counters are known to be zero (its owner
                                   // cleared them once; every launch leaves them zero)
    bool decoded = false;  // records, cull boxes and cleared counters of these B x N splats are in the
                           // workspace already (launch_propose): skip the decode launch
    PeerStores peers;
};
int evaluate(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
             float k_sigma, const float bg[3], const float *d_target, const float *d_mask,
             int mode, float beta, float *d_fitness, void *d_images, int image_u8,
             void *d_workspace, size_t workspace_bytes_given, cudaStream_t stream,
             const EvalOptions &opt = EvalOptions());
// Largest split (1, 2, 4, 8) that keeps B * tiles * split CTAs within half a wave of the device
// and every genome segment worth a CTA, 1 for deep genomes; the policy behind split = 0.
int choose_split(int B, int N, int H, int W);

}  // namespace ggs
