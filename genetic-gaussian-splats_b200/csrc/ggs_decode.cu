// Genome decode for the whole population in one launch (arithmetic: ggs_decode_math.cuh).
#include "ggs_decode_math.cuh"

namespace ggs {
namespace {

// Stage this block's rows into shared memory with 128-bit coalesced loads, then hand each
// thread its own row (stride `cols` floats: conflict-free for the usual cols = 9).
__device__ __forceinline__ const float *stage_rows(const float *__restrict__ genomes, int cols,
                                                   int64_t rows, float *sm, float *tmp)
{
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * kDecodeThreads;
    const int nrows = (int)min((int64_t)kDecodeThreads, rows - row0);
    const float *src = genomes + row0 * cols;
    if (cols <= kDecodeStageMaxCols) {
        const int nfl = nrows * cols;
        int done = 0;
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            const int n4 = nfl >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(sm);
            for (int i = tid; i < n4; i += kDecodeThreads) dst4[i] = __ldg(src4 + i);
            done = n4 << 2;
        }
        for (int i = done + tid; i < nfl; i += kDecodeThreads) sm[i] = __ldg(src + i);
        __syncthreads();
        return (tid < nrows) ? sm + tid * cols : nullptr;
    }
    if (tid >= nrows) return nullptr;
#pragma unroll
    for (int k = 0; k < 9; ++k) tmp[k] = __ldg(src + (int64_t)tid * cols + k);
    return tmp;
}

template <bool kAxes>
__global__ void __launch_bounds__(kDecodeThreads)
decode_kernel(const float *__restrict__ genomes, int cols, int64_t rows, int H, int W,
              float k_sigma, float4 *__restrict__ rec, uint2 *__restrict__ aabb,
              float *__restrict__ raw_f, int32_t *__restrict__ raw_i, int *__restrict__ counters,
              int n_counters)
{
    extern __shared__ __align__(16) float sm[];
    pdl_wait();
    pdl_trigger();
    if (blockIdx.x == 0 && counters != nullptr)
        for (int i = threadIdx.x; i < n_counters; i += kDecodeThreads) counters[i] = 0;

    float tmp[9];
    const float *g = stage_rows(genomes, cols, rows, sm, tmp);
    if (g == nullptr) return;
    const int64_t row = (int64_t)blockIdx.x * kDecodeThreads + threadIdx.x;

    const Chol ch = kAxes ? encode_axes(g) : load_chol(g);
    const Decoded d = decode_chol(ch, H, W, k_sigma);

    if (raw_f != nullptr) {
        raw_f[0 * rows + row] = d.cx;
        raw_f[1 * rows + row] = d.cy;
        raw_f[2 * rows + row] = d.sxx;
        raw_f[3 * rows + row] = d.sxy;
        raw_f[4 * rows + row] = d.syy;
        raw_f[5 * rows + row] = d.rc;
        raw_f[6 * rows + row] = d.gc;
        raw_f[7 * rows + row] = d.bc;
        raw_f[8 * rows + row] = d.a;
        raw_i[0 * rows + row] = d.x0;
        raw_i[1 * rows + row] = d.x1;
        raw_i[2 * rows + row] = d.y0;
        raw_i[3 * rows + row] = d.y1;
    }
    if (rec != nullptr) {
        SplatRec r;
        uint2 box;
        make_record(d, r, box);
        const float4 *rv = reinterpret_cast<const float4 *>(&r);
        rec[row * 3 + 0] = rv[0];
        rec[row * 3 + 1] = rv[1];
        rec[row * 3 + 2] = rv[2];
        aabb[row] = box;
    }
}

__global__ void __launch_bounds__(kDecodeThreads)
encode_kernel(const float *__restrict__ axes, int cols, int64_t rows, float *__restrict__ chol)
{
    extern __shared__ __align__(16) float sm[];
    pdl_wait();
    pdl_trigger();
    float tmp[9];
    const float *g = stage_rows(axes, cols, rows, sm, tmp);
    if (g == nullptr) return;
    const int64_t row = (int64_t)blockIdx.x * kDecodeThreads + threadIdx.x;
    const Chol c = encode_axes(g);
    float *o = chol + row * 9;
    o[0] = c.x;
    o[1] = c.y;
    o[2] = c.log_l11;
    o[3] = c.log_l22;
    o[4] = c.l21;
    o[5] = c.r;
    o[6] = c.g;
    o[7] = c.b;
    o[8] = c.a;
}

inline size_t stage_bytes(int cols)
{
    return cols <= kDecodeStageMaxCols ? (size_t)kDecodeThreads * cols * sizeof(float) : 16;
}

}  // namespace

cudaError_t launch_decode(const float *d_genomes, int layout, int64_t rows, int cols, int H, int W,
                          float k_sigma, float4 *rec, uint2 *aabb, float *raw_f, int32_t *raw_i,
                          int *counters, int n_counters, cudaStream_t stream)
{
    if (rows <= 0) {
        // no splats: the raster still needs its per-candidate counters cleared
        if (counters != nullptr && n_counters > 0)
            return cudaMemsetAsync(counters, 0, (size_t)n_counters * sizeof(int), stream);
        return cudaSuccess;
    }
    const unsigned grid = (unsigned)((rows + kDecodeThreads - 1) / kDecodeThreads);
    const size_t smem = stage_bytes(cols);
    if (layout == GGS_LAYOUT_AXES_ANGLE)
        return launch_kernel(decode_kernel<true>, grid, kDecodeThreads, smem, stream, d_genomes, cols,
                             rows, H, W, k_sigma, rec, aabb, raw_f, raw_i, counters, n_counters);
    return launch_kernel(decode_kernel<false>, grid, kDecodeThreads, smem, stream, d_genomes, cols,
                         rows, H, W, k_sigma, rec, aabb, raw_f, raw_i, counters, n_counters);
}

cudaError_t launch_encode(const float *d_axes, int64_t rows, int cols, float *d_chol,
                          cudaStream_t stream)
{
    if (rows <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((rows + kDecodeThreads - 1) / kDecodeThreads);
    return launch_kernel(encode_kernel, grid, kDecodeThreads, stage_bytes(cols), stream, d_axes, cols,
                         rows, d_chol);
}

}  // namespace ggs
