// Hardware probes for the roofline denominators.  MEASURED_PEAKS.json carries HBM and bf16
// tensor peaks only; this path is bound by the FP32 FMA pipe, so
// bench.py measures that peak itself: dependent-chain-free FFMA, packed FFMA2 (fma.rn.f32x2,
// new on sm_100) and MUFU.EX2 issue rates, timed with CUDA events.
#include <stdio.h>

#include "ggs_common.cuh"

namespace ggs {
namespace {

constexpr int kProbeThreads = 256;
constexpr int kChains = 16;  // independent accumulators per thread

__global__ void __launch_bounds__(kProbeThreads) ffma_kernel(float *out, int iters, float a, float b)
{
    float acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == 12345.678f) out[0] = s;  // never true; keeps the chain alive
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void __launch_bounds__(kProbeThreads)
ffma2_kernel(float *out, int iters, unsigned long long a, unsigned long long b)
{
    unsigned long long acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = (unsigned long long)(threadIdx.x + k) * 0x3f8000003f800000ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) acc[k] = ffma2(acc[k], a, b);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s ^= acc[k];
    if (s == 0x123456789abcdefull) out[0] = 1.0f;
}

__global__ void __launch_bounds__(kProbeThreads) mufu_kernel(float *out, int iters, float seed)
{
    float acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = seed * (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kChains; ++k)
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[k]));
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == 12345.678f) out[0] = s;
}

// Packed and scalar FMA chains interleaved (kP packed + kS scalar per thread): shows whether
// FFMA2 (fmaheavy) and scalar FFMA (fmaheavy + fmalite) add up beyond either alone.
template <int kP, int kS>
__global__ void __launch_bounds__(kProbeThreads)
mix_kernel(float *out, int iters, unsigned long long a2, unsigned long long b2, float a, float b)
{
    unsigned long long p[kP];
    float q[kS];
#pragma unroll
    for (int k = 0; k < kP; ++k) p[k] = (unsigned long long)(threadIdx.x + k) * 0x3f8000003f800000ull;
#pragma unroll
    for (int k = 0; k < kS; ++k) q[k] = (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < (kP > kS ? kP : kS); ++k) {
            if (k < kP) p[k] = ffma2(p[k], a2, b2);
            if (k < kS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(q[k]) : "f"(a), "f"(b));
        }
    }
    unsigned long long s = 0;
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < kP; ++k) s ^= p[k];
#pragma unroll
    for (int k = 0; k < kS; ++k) t += q[k];
    if (s == 0x123456789abcdefull || t == 12345.678f) out[0] = 1.0f;
}

template <typename F>
cudaError_t time_ms(F launch, float *ms)
{
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    launch();  // warm-up
    launch();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return e;
        float t = 0.0f;
        cudaEventElapsedTime(&t, e0, e1);
        best = t < best ? t : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = best;
    return cudaGetLastError();
}

}  // namespace

static float g_mix[3];

cudaError_t probe_peaks(float *h_out5)
{
    int dev = 0, sms = 0, khz = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev)) != cudaSuccess) return e;
    float *d_out = nullptr;
    if ((e = cudaMalloc(&d_out, 16)) != cudaSuccess) return e;

    const int blocks = sms * 8;  // 2048 threads per SM: full occupancy
    const int iters = 4096;
    const double lanes = (double)blocks * kProbeThreads * kChains * (double)iters;
    float ms = 0.0f;

    e = time_ms([&] { ffma_kernel<<<blocks, kProbeThreads>>>(d_out, iters, 0.999f, 0.001f); }, &ms);
    if (e != cudaSuccess) return e;
    h_out5[0] = (float)(2.0 * lanes / (ms * 1e-3) / 1e12);

    e = time_ms([&] { ffma2_kernel<<<blocks, kProbeThreads>>>(d_out, iters, 0x3f7fbe773f7fbe77ull,
                                                             0x3a83126f3a83126full); }, &ms);
    if (e != cudaSuccess) return e;
    h_out5[1] = (float)(4.0 * lanes / (ms * 1e-3) / 1e12);

    e = time_ms([&] { mufu_kernel<<<blocks, kProbeThreads>>>(d_out, iters / 4, -0.001f); }, &ms);
    if (e != cudaSuccess) return e;
    h_out5[2] = (float)(lanes / 4.0 / (ms * 1e-3) / 1e9);

    // mixed packed + scalar: reported as FMA-lane TFLOP/s (packed counts 4 flop, scalar 2)
    e = time_ms([&] { mix_kernel<8, 8><<<blocks, kProbeThreads>>>(d_out, iters, 0x3f7fbe773f7fbe77ull,
                                                                  0x3a83126f3a83126full, 0.999f, 0.001f); }, &ms);
    if (e != cudaSuccess) return e;
    g_mix[0] = (float)((double)blocks * kProbeThreads * iters * (8 * 4.0 + 8 * 2.0) / (ms * 1e-3) / 1e12);
    e = time_ms([&] { mix_kernel<12, 4><<<blocks, kProbeThreads>>>(d_out, iters, 0x3f7fbe773f7fbe77ull,
                                                                   0x3a83126f3a83126full, 0.999f, 0.001f); }, &ms);
    if (e != cudaSuccess) return e;
    g_mix[1] = (float)((double)blocks * kProbeThreads * iters * (12 * 4.0 + 4 * 2.0) / (ms * 1e-3) / 1e12);
    e = time_ms([&] { mix_kernel<8, 4><<<blocks, kProbeThreads>>>(d_out, iters, 0x3f7fbe773f7fbe77ull,
                                                                  0x3a83126f3a83126full, 0.999f, 0.001f); }, &ms);
    if (e != cudaSuccess) return e;
    g_mix[2] = (float)((double)blocks * kProbeThreads * iters * (8 * 4.0 + 4 * 2.0) / (ms * 1e-3) / 1e12);
    fprintf(stderr, "[ggs probe] mixed FFMA2+FFMA TFLOP/s: 8p+8s %.1f, 12p+4s %.1f, 8p+4s %.1f\n",
            g_mix[0], g_mix[1], g_mix[2]);

    h_out5[3] = (float)sms;
    h_out5[4] = (float)khz / 1000.0f;
    cudaFree(d_out);
    return cudaSuccess;
}

}  // namespace ggs
