// The fitness all-gather of a population sharded over the GPUs of one box, WITHOUT a collective
// launch (the reference has no multi-GPU code at all).
//
// Every rank (one process per GPU) owns one exported device allocation
//     [ arrival flags: one uint32 per sender | status | gathered vector, two epochs deep ]
// and maps the allocations of the other ranks through CUDA IPC (NVLink / NVSwitch peer access).
// The raster kernel's candidate-finishing CTA stores the fitness value straight into EVERY rank's
// gathered vector (peer_publish, ggs_common.cuh: st.relaxed.sys over NVLink, 4 bytes per
// candidate per peer), and the CTA that finishes the launch's last candidate raises this rank's
// flag on every rank with a system-scope release.  A consumer waits for `world` flags with
// acquire loads (peer_wait_kernel here, or the head of the GA engine's select kernel).  The
// exchange is therefore fused into the compute kernel: no NCCL launch, no extra pass over the
// data, and the stores of early candidates overlap the rendering of late ones.
//
// Epochs: every gather has a number (1, 2, ...) that all ranks advance in lockstep; epoch e uses
// half (e & 1) of the gathered buffer, so a rank that runs one step ahead never overwrites values
// a slower rank is still reading (it cannot run two ahead: its next step needs that rank's flag).
#include <string.h>

#include <new>

#include "ggs_common.cuh"

namespace ggs {
namespace {

constexpr size_t kFlagBytes = 256;    // kMaxPeers flags, padded
constexpr size_t kStatusBytes = 256;  // [0] = 1 after a wait timed out

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

namespace ggs {
namespace {

__global__ void __launch_bounds__(32) peer_wait_kernel(const unsigned *flags, int world, unsigned epoch,
                                                       int *status)
{
    pdl_wait();
    pdl_trigger();
    peer_wait(flags, world, epoch, status);
}

// A rank whose shard is empty has no raster launch to raise its flag.
__global__ void __launch_bounds__(32) peer_signal_kernel(PeerStores p)
{
    pdl_wait();
    pdl_trigger();
    if ((int)threadIdx.x < p.n)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag[threadIdx.x] + p.rank), "r"(p.epoch)
                     : "memory");
}

}  // namespace
}  // namespace ggs

using namespace ggs;

struct ggs_peers {
    int device = 0, rank = 0, world = 1, capacity = 0;
    char *own = nullptr;                 // this rank's exported allocation
    char *base[kMaxPeers] = {};          // every rank's allocation as mapped here (own for rank)
    bool opened[kMaxPeers] = {};
    int *done = nullptr;                 // local: candidates published by the launch in flight
    unsigned epoch = 0;                  // last epoch handed out
    bool connected = false;
};

namespace ggs {

unsigned *peers_flags(ggs_peers *p) { return reinterpret_cast<unsigned *>(p->own); }
int *peers_status(ggs_peers *p) { return reinterpret_cast<int *>(p->own + kFlagBytes); }
float *peers_gathered(ggs_peers *p, unsigned epoch)
{
    return reinterpret_cast<float *>(p->own + kFlagBytes + kStatusBytes) + (size_t)(epoch & 1u) * p->capacity;
}
int peers_rank(const ggs_peers *p) { return p->rank; }
int peers_world(const ggs_peers *p) { return p->world; }
int peers_capacity(const ggs_peers *p) { return p->capacity; }
bool peers_ready(const ggs_peers *p) { return p->connected; }

// The stores of the next gather: epoch number taken, pointers of every rank's buffer half.
PeerStores peers_next(ggs_peers *p, int offset)
{
    PeerStores s;
    p->epoch += 1;
    s.n = p->world;
    s.rank = p->rank;
    s.offset = offset;
    s.epoch = p->epoch;
    for (int r = 0; r < p->world; ++r) {
        s.flag[r] = reinterpret_cast<unsigned *>(p->base[r]);
        s.fit[r] = reinterpret_cast<float *>(p->base[r] + kFlagBytes + kStatusBytes) +
                   (size_t)(p->epoch & 1u) * p->capacity;
    }
    s.done = p->done;
    return s;
}

cudaError_t peers_signal_empty(const PeerStores &s, cudaStream_t st)
{
    return launch_kernel(peer_signal_kernel, 1, 32, 0, st, s);
}

cudaError_t peers_wait(ggs_peers *p, unsigned epoch, cudaStream_t st)
{
    return launch_kernel(peer_wait_kernel, 1, 32, 0, st, (const unsigned *)peers_flags(p), p->world, epoch,
                         peers_status(p));
}

}  // namespace ggs

extern "C" {

int ggs_peers_create(int device, int rank, int world, int capacity, ggs_peers **out)
{
    if (out == nullptr) {
        set_error("ggs_peers_create: out is NULL");
        return GGS_EINVAL;
    }
    *out = nullptr;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || capacity < 1) {
        set_error("ggs_peers_create: bad arguments (rank %d of %d, at most %d ranks; capacity %d)", rank,
                  world, kMaxPeers, capacity);
        return GGS_EINVAL;
    }
    int n = 0;
    GGS_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        set_error("ggs_peers_create: device %d not visible (%d devices)", device, n);
        return GGS_ENODEVICE;
    }
    DeviceGuard on_device(device);
    GGS_TRY(on_device.status());
    ggs_peers *p = new (std::nothrow) ggs_peers();
    if (!p) {
        set_error("out of host memory");
        return GGS_EINVAL;
    }
    p->device = device;
    p->rank = rank;
    p->world = world;
    p->capacity = (capacity + 63) / 64 * 64;
    const size_t bytes = kFlagBytes + kStatusBytes + 2 * (size_t)p->capacity * sizeof(float);
    cudaError_t e = cudaMalloc(&p->own, bytes);
    if (e == cudaSuccess) e = cudaMemset(p->own, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->done, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(p->done, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        ggs_peers_destroy(p);
        return fail(e, "ggs_peers_create: cudaMalloc");
    }
    p->base[rank] = p->own;
    p->connected = (world == 1);
    *out = p;
    return GGS_OK;
}

void ggs_peers_destroy(ggs_peers *p)
{
    if (!p) return;
    DeviceGuard on_device(p->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->base[r]);
    if (p->own) cudaFree(p->own);
    if (p->done) cudaFree(p->done);
    delete p;
}

int ggs_peers_export(ggs_peers *p, void *h_handle)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == GGS_IPC_HANDLE_BYTES, "IPC handle size");
    if (!p || !h_handle) {
        set_error("ggs_peers_export: NULL argument");
        return GGS_EINVAL;
    }
    DeviceGuard on_device(p->device);
    GGS_TRY(on_device.status());
    cudaIpcMemHandle_t h;
    GGS_TRY(cudaIpcGetMemHandle(&h, p->own));
    memcpy(h_handle, &h, sizeof(h));
    return GGS_OK;
}

int ggs_peers_connect(ggs_peers *p, const void *h_handles)
{
    if (!p || !h_handles) {
        set_error("ggs_peers_connect: NULL argument");
        return GGS_EINVAL;
    }
    DeviceGuard on_device(p->device);
    GGS_TRY(on_device.status());
    const char *src = static_cast<const char *>(h_handles);
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank || p->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, src + (size_t)r * GGS_IPC_HANDLE_BYTES, sizeof(h));
        void *mapped = nullptr;
        GGS_TRY(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
        p->base[r] = static_cast<char *>(mapped);
        p->opened[r] = true;
    }
    p->connected = true;
    return GGS_OK;
}

int ggs_peers_connect_local(ggs_peers *p, ggs_peers *const *all)
{
    if (!p || !all) {
        set_error("ggs_peers_connect_local: NULL argument");
        return GGS_EINVAL;
    }
    DeviceGuard on_device(p->device);
    GGS_TRY(on_device.status());
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        if (!all[r] || all[r]->world != p->world || all[r]->rank != r || all[r]->capacity != p->capacity) {
            set_error("ggs_peers_connect_local: entry %d is not rank %d of the same group", r, r);
            return GGS_EINVAL;
        }
        if (all[r]->device != p->device) {
            int can = 0;
            GGS_TRY(cudaDeviceCanAccessPeer(&can, p->device, all[r]->device));
            if (!can) {
                set_error("ggs_peers_connect_local: device %d cannot access device %d", p->device, all[r]->device);
                return GGS_ENODEVICE;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(e, "cudaDeviceEnablePeerAccess");
            (void)cudaGetLastError();
        }
        p->base[r] = all[r]->own;
    }
    p->connected = true;
    return GGS_OK;
}

int ggs_fitness_allgather(ggs_peers *p, const float *d_genomes, int layout, int B, int N, int cols,
                          int H, int W, float k_sigma, const float *d_target, const float *d_mask,
                          int mode, float boost_beta, int offset, int total, void *d_workspace,
                          size_t workspace_bytes_given, const float **d_gathered, void *stream)
{
    if (!p || !p->connected) {
        set_error("ggs_fitness_allgather: peers not connected");
        return GGS_EINVAL;
    }
    if (B < 0 || offset < 0 || total < 1 || offset + B > total || total > p->capacity) {
        set_error("ggs_fitness_allgather: shard [%d, %d) of %d does not fit the gathered vector (%d)",
                  offset, offset + B, total, p->capacity);
        return GGS_EINVAL;
    }
    if (mode < GGS_MODE_PLAIN || mode > GGS_MODE_BOOST || (mode != GGS_MODE_PLAIN && !d_mask) ||
        (B > 0 && (!d_genomes || !d_target)) || cols < 9 || N < 0) {
        set_error("ggs_fitness_allgather: bad arguments");
        return GGS_EINVAL;
    }
    if ((layout != GGS_LAYOUT_AXES_ANGLE && layout != GGS_LAYOUT_CHOLESKY) || H < 1 || W < 1 ||
        H > GGS_MAX_SIDE || W > GGS_MAX_SIDE) {
        set_error("ggs_fitness_allgather: bad layout or image size");
        return GGS_EINVAL;
    }
    // everything that can fail is checked BEFORE an epoch is taken: a rank that takes one and then
    // does not deliver would leave its peers waiting for the time-out
    if (B > 0 && (d_workspace == nullptr || workspace_bytes_given < workspace_bytes(B, N, H, W) ||
                  (reinterpret_cast<uintptr_t>(d_workspace) & 255u) != 0)) {
        set_error("ggs_fitness_allgather: workspace missing, misaligned or smaller than %zu bytes",
                  workspace_bytes(B, N, H, W));
        return GGS_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DeviceGuard on_device(p->device);
    GGS_TRY(on_device.status());
    EvalOptions opt;
    opt.split = choose_split(total, N, H, W);  // the whole population's configuration on every rank
    opt.peers = peers_next(p, offset);
    const unsigned epoch = opt.peers.epoch;
    float *mine = peers_gathered(p, epoch);
    if (B > 0) {
        const float white[3] = {1.0f, 1.0f, 1.0f};
        int rc = evaluate(d_genomes, layout, B, N, cols, H, W, k_sigma, white, d_target, d_mask, mode,
                          boost_beta, mine + offset, nullptr, 0, d_workspace, workspace_bytes_given, st, opt);
        if (rc) return rc;
    } else {
        GGS_TRY(peers_signal_empty(opt.peers, st));
    }
    GGS_TRY(peers_wait(p, epoch, st));
    if (d_gathered) *d_gathered = mine;
    return GGS_OK;
}

int ggs_peers_status(ggs_peers *p, void *stream)
{
    if (!p) {
        set_error("ggs_peers_status: NULL argument");
        return GGS_EINVAL;
    }
    DeviceGuard on_device(p->device);
    GGS_TRY(on_device.status());
    int status = 0;
    GGS_TRY(cudaMemcpyAsync(&status, peers_status(p), sizeof(int), cudaMemcpyDeviceToHost,
                            static_cast<cudaStream_t>(stream)));
    GGS_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (status != 0) {
        set_error("a peer wait timed out: a rank did not deliver its fitness values");
        return GGS_ECUDA;
    }
    return GGS_OK;
}

}  // extern "C"
