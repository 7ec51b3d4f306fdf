// Speed-of-light of the raster kernel's composite loop, layer by layer (B200, sm_100a).
//
// The composite loop of ggs_raster.cu runs at ~79 cycles per (warp, list entry) per scheduler
// with the FMA pipe 66 % and the issue slots 77 % busy, and removing instructions from its
// prologue does not make it faster.  This microbenchmark rebuilds the loop body from the inside
// out -- same named-register PTX, same operand pattern, same occupancy (8 CTAs x 128 threads per
// SM, 64 registers) -- and times each layer on its own:
//   M0  the 20 packed blend operations of 4 row pairs (FMUL2, 3 x FFMA2, FADD2 each)
//   M1  + the 5 packed recurrence steps (F *= G, G *= H)
//   M2  + exponents: 3 FFMA2 + 2 scalar + 4 MUFU.EX2
//   M3  + the per-entry prologue: 3 x LDS.128 from a uniform address, 6 scalar FP, lane-mask select,
//         band byte test, loop control (what the real loop does on its recurrence path)
// and a few alternative formulations of M0.  Output: cycles per iteration per scheduler (SMSP)
// with 8 warps resident, i.e. directly comparable with the kernel's 79.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench_blend tools/microbench_blend.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

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
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#define PX_DECLARE() asm volatile(".reg .b64 ggs_r<4>, ggs_g<4>, ggs_b<4>, ggs_t<4>, ggs_F, ggs_G;")
#define PX_INIT(k)                                                                         \
    asm volatile("mov.b64 ggs_r" #k ", 0;\n\tmov.b64 ggs_g" #k ", 0;\n\tmov.b64 ggs_b" #k   \
                 ", 0;\n\tmov.b64 ggs_t" #k ", {%0, %1};" ::"f"(1.0f), "f"(1.0f));
#define SET_F(f0_, f1_) asm volatile("mov.b64 ggs_F, {%0, %1};" ::"f"(f0_), "f"(f1_));
#define SET_G(g0_, g1_) asm volatile("mov.b64 ggs_G, {%0, %1};" ::"f"(g0_), "f"(g1_));
#define STEP_F() asm volatile("mul.rn.f32x2 ggs_F, ggs_F, ggs_G;");
#define STEP_G(H_) asm volatile("mul.rn.f32x2 ggs_G, ggs_G, %0;" ::"l"(H_));
// the kernel's blend: W = F*T, C += W*col, T -= W
#define BLEND(k)                                                                          \
    asm volatile("{\n\t.reg .b64 w;\n\t"                                                 \
                 "mul.rn.f32x2 w, ggs_F, ggs_t" #k ";\n\t"                                \
                 "fma.rn.f32x2 ggs_r" #k ", w, %0, ggs_r" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_g" #k ", w, %1, ggs_g" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_b" #k ", w, %2, ggs_b" #k ";\n\t"                      \
                 "sub.rn.f32x2 ggs_t" #k ", ggs_t" #k ", w;\n\t}" ::"l"(R2), "l"(G2), "l"(B2));
// alternative: T updated by an FMA that does not wait for W (T = -F*T + T)
#define BLEND_TFMA(k)                                                                     \
    asm volatile("{\n\t.reg .b64 w, nf;\n\t"                                             \
                 "mul.rn.f32x2 w, ggs_F, ggs_t" #k ";\n\t"                                \
                 "xor.b64 nf, ggs_F, 0x8000000080000000;\n\t"                                               \
                 "fma.rn.f32x2 ggs_t" #k ", nf, ggs_t" #k ", ggs_t" #k ";\n\t"            \
                 "fma.rn.f32x2 ggs_r" #k ", w, %0, ggs_r" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_g" #k ", w, %1, ggs_g" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_b" #k ", w, %2, ggs_b" #k ";\n\t}" ::"l"(R2), "l"(G2), "l"(B2));
// alternative: the blue channel as two scalar FFMAs (same pipe cycles, one more issue slot)
#define BLEND_SCALARB(k)                                                                  \
    asm volatile("{\n\t.reg .b64 w;\n\t.reg .f32 w0, w1, b0, b1;\n\t"                     \
                 "mul.rn.f32x2 w, ggs_F, ggs_t" #k ";\n\t"                                \
                 "fma.rn.f32x2 ggs_r" #k ", w, %0, ggs_r" #k ";\n\t"                      \
                 "fma.rn.f32x2 ggs_g" #k ", w, %1, ggs_g" #k ";\n\t"                      \
                 "mov.b64 {w0, w1}, w;\n\tmov.b64 {b0, b1}, ggs_b" #k ";\n\t"             \
                 "fma.rn.f32 b0, w0, %3, b0;\n\tfma.rn.f32 b1, w1, %3, b1;\n\t"           \
                 "mov.b64 ggs_b" #k ", {b0, b1};\n\t"                                     \
                 "sub.rn.f32x2 ggs_t" #k ", ggs_t" #k ", w;\n\t}" ::"l"(R2), "l"(G2), "l"(B2), "f"(Bs));
#define READ(k, acc)                                                                       \
    {                                                                                      \
        float a0, a1, a2, a3, a4, a5, a6, a7;                                              \
        asm volatile("mov.b64 {%0, %1}, ggs_r" #k ";\n\tmov.b64 {%2, %3}, ggs_g" #k          \
                     ";\n\tmov.b64 {%4, %5}, ggs_b" #k ";\n\tmov.b64 {%6, %7}, ggs_t" #k ";" \
                     : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7)); \
        acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;                                      \
    }
#define PAIRS(M) M(0) M(1) M(2) M(3)

constexpr int kThreads = 128;
constexpr int kEntries = 256;

// mode: 0 = M0, 1 = M1, 2 = M2, 3 = M3, 10 = M0 with BLEND_TFMA, 11 = M0 with BLEND_SCALARB,
//       12 = M0 with pairs interleaved op by op (ILP across pairs)
template <int kMode>
__global__ void __launch_bounds__(kThreads, 8) loop_kernel(const float4 *__restrict__ entries, float *out, int iters)
{
    __shared__ float4 s_list[kEntries * 3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kEntries * 3; i += kThreads) s_list[i] = entries[i];
    __syncthreads();
    const float Xf = __shfl_sync(0xffffffffu, (float)lane, lane);
    const float Ybf = __shfl_sync(0xffffffffu, (float)(warp * 8), lane);
    const unsigned lanebit = __shfl_sync(0xffffffffu, 1u << lane, lane);
    const unsigned band_sel = __shfl_sync(0xffffffffu, 0x4440u + (unsigned)warp, lane);
    PX_DECLARE();
    PAIRS(PX_INIT)
    const float4 e0 = s_list[0], e1 = s_list[1], e2 = s_list[2];
    f2_t R2 = bcast2(e1.z), G2 = bcast2(e1.w), B2 = bcast2(e2.x);
    if (kMode == 13) {  // genuine register pairs (c, c) instead of one register broadcast (Rn.F32)
        const float z = e0.x * 0.0f;  // zero, but not provably
        R2 = pack2(e1.z, e1.z + z), G2 = pack2(e1.w, e1.w + z), B2 = pack2(e2.x, e2.x + z);
    }
    float Bs = e2.x;
    (void)Bs;
    f2_t H2 = bcast2(e2.w);
    SET_F(1e-7f * Xf, 1.1e-7f * Xf)   // per-lane values: vector registers, as in the kernel
    SET_G(1.0f + 1e-7f * Xf, 1.0f)
    for (int rep = 0; rep < iters; ++rep) {
        if (kMode == 3) {
            const float4 *q = s_list;
#pragma unroll 1
            for (int s = 0; s < kEntries; ++s, q += 3) {
                const float4 q2 = q[2];
                if ((s & 7) == 7) {
                    float tmax = 0.0f;
#define TMAX(k)                                                                      \
    {                                                                                \
        float ta, tb;                                                                \
        asm volatile("mov.b64 {%0, %1}, ggs_t" #k ";" : "=f"(ta), "=f"(tb));         \
        tmax = fmaxf(tmax, fmaxf(ta, tb));                                           \
    }
                    PAIRS(TMAX)
                    if (__all_sync(0xffffffffu, tmax < 2.384185791015625e-07f)) break;
                }
                const unsigned c = __byte_perm(__float_as_uint(q2.z), 0u, band_sel);
                if (c == 0x0fu) continue;
                const float4 q0 = q[0];
                const float4 q1 = q[1];
                const bool in_x = (__float_as_uint(q2.y) & lanebit) != 0u;
                const float qx = Xf - q0.x;
                const float t1 = q0.w * qx;
                float t0 = fmaf(q0.z * qx, qx, q1.y);
                t0 = in_x ? t0 : -INFINITY;
                const float dy = Ybf - q0.y;
                const f2_t QY = pack2(dy, dy + 1.0f);
                const f2_t CQ2 = bcast2(q1.x), T12 = bcast2(t1), T02 = bcast2(t0);
                R2 = bcast2(q1.z), G2 = bcast2(q1.w), B2 = bcast2(q2.x);
                if (c == 0x70u) {
                    const f2_t E = fma2(fma2(CQ2, QY, T12), QY, T02);
                    const float c4 = 4.0f * q1.x;
                    const f2_t D = fma2(bcast2(c4), QY, bcast2(fmaf(2.0f, t1, c4)));
                    float ea, eb, da, db;
                    unpack2(E, ea, eb);
                    unpack2(D, da, db);
                    SET_F(ex2_approx(ea), ex2_approx(eb))
                    SET_G(ex2_approx(da), ex2_approx(db))
                    H2 = bcast2(q2.w);
                    BLEND(0) STEP_F() STEP_G(H2) BLEND(1) STEP_F() STEP_G(H2) BLEND(2) STEP_F() BLEND(3)
                }
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < kEntries; ++s) {
                if (kMode == 2) {
                    // exponents from registers that change every iteration (no hoisting)
                    const float t1 = __uint_as_float(__float_as_uint(e0.w) + (unsigned)s);
                    const float t0 = e1.y, dy = e0.y;
                    const f2_t QY = pack2(dy, dy + 1.0f);
                    const f2_t CQ2 = bcast2(e1.x), T12 = bcast2(t1), T02 = bcast2(t0);
                    const f2_t E = fma2(fma2(CQ2, QY, T12), QY, T02);
                    const float c4 = 4.0f * e1.x;
                    const f2_t D = fma2(bcast2(c4), QY, bcast2(fmaf(2.0f, t1, c4)));
                    float ea, eb, da, db;
                    unpack2(E, ea, eb);
                    unpack2(D, da, db);
                    SET_F(ex2_approx(ea), ex2_approx(eb))
                    SET_G(ex2_approx(da), ex2_approx(db))
                }
                if (kMode == 0 || kMode == 13) {
                    BLEND(0) BLEND(1) BLEND(2) BLEND(3)
                } else if (kMode == 10) {
                    BLEND_TFMA(0) BLEND_TFMA(1) BLEND_TFMA(2) BLEND_TFMA(3)
                } else if (kMode == 11) {
                    BLEND_SCALARB(0) BLEND_SCALARB(1) BLEND_SCALARB(2) BLEND_SCALARB(3)
                } else if (kMode == 12) {
                    asm volatile("{\n\t.reg .b64 w<4>;\n\t"
                                 "mul.rn.f32x2 w0, ggs_F, ggs_t0;\n\tmul.rn.f32x2 w1, ggs_F, ggs_t1;\n\t"
                                 "mul.rn.f32x2 w2, ggs_F, ggs_t2;\n\tmul.rn.f32x2 w3, ggs_F, ggs_t3;\n\t"
                                 "fma.rn.f32x2 ggs_r0, w0, %0, ggs_r0;\n\tfma.rn.f32x2 ggs_r1, w1, %0, ggs_r1;\n\t"
                                 "fma.rn.f32x2 ggs_r2, w2, %0, ggs_r2;\n\tfma.rn.f32x2 ggs_r3, w3, %0, ggs_r3;\n\t"
                                 "fma.rn.f32x2 ggs_g0, w0, %1, ggs_g0;\n\tfma.rn.f32x2 ggs_g1, w1, %1, ggs_g1;\n\t"
                                 "fma.rn.f32x2 ggs_g2, w2, %1, ggs_g2;\n\tfma.rn.f32x2 ggs_g3, w3, %1, ggs_g3;\n\t"
                                 "fma.rn.f32x2 ggs_b0, w0, %2, ggs_b0;\n\tfma.rn.f32x2 ggs_b1, w1, %2, ggs_b1;\n\t"
                                 "fma.rn.f32x2 ggs_b2, w2, %2, ggs_b2;\n\tfma.rn.f32x2 ggs_b3, w3, %2, ggs_b3;\n\t"
                                 "sub.rn.f32x2 ggs_t0, ggs_t0, w0;\n\tsub.rn.f32x2 ggs_t1, ggs_t1, w1;\n\t"
                                 "sub.rn.f32x2 ggs_t2, ggs_t2, w2;\n\tsub.rn.f32x2 ggs_t3, ggs_t3, w3;\n\t}" ::"l"(R2),
                                 "l"(G2), "l"(B2));
                } else {
                    BLEND(0) STEP_F() STEP_G(H2) BLEND(1) STEP_F() STEP_G(H2) BLEND(2) STEP_F() BLEND(3)
                    if (kMode == 1) SET_F(1e-7f * Xf, 1.1e-7f * Xf)
                }
            }
        }
    }
    float acc = 0.0f;

    READ(0, acc) READ(1, acc) READ(2, acc) READ(3, acc)
    if (acc == 12345.678f) out[0] = acc;
}

template <int kMode>
static double run(const float4 *d_entries, float *d_out, int iters, int sms, double ghz, const char *name, int packed,
                  int scalar)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sms * 8;
    loop_kernel<kMode><<<blocks, kThreads>>>(d_entries, d_out, 2);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        loop_kernel<kMode><<<blocks, kThreads>>>(d_entries, d_out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        printf("%s: %s\n", name, cudaGetErrorString(e));
        return 0;
    }
    // 8 warps per scheduler, each doing iters * kEntries loop iterations
    const double cyc = best * 1e-3 * ghz * 1e9 / ((double)iters * kEntries * 8.0);
    printf("%-64s %7.2f cycles / (warp, entry) / scheduler", name, cyc);
    if (packed) printf("   [%d packed + %d scalar FMA-pipe ops = %d pipe cycles nominal]", packed, scalar, 2 * packed + scalar);
    printf("\n");
    return cyc;
}

int main()
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("device: %d SMs, %.3f GHz (cudaDevAttrClockRate); 8 CTAs x 128 threads per SM\n", sms, ghz);
    // list entries that all take the recurrence path: every band fully covered, every lane inside
    float4 *h = (float4 *)malloc(sizeof(float4) * kEntries * 3);
    for (int i = 0; i < kEntries; ++i) {
        const float Cq = -1e-4f;
        h[3 * i + 0] = make_float4(16.0f, 16.0f, -1e-4f, 1e-5f);             // cx, cy, A, Bq
        h[3 * i + 1] = make_float4(Cq, -20.0f, 0.3f, 0.5f);                  // Cq, log2 alpha (tiny alpha), r, g
        unsigned mask = 0xffffffffu, code = 0x70707070u;
        float fm, fc;
        memcpy(&fm, &mask, 4);
        memcpy(&fc, &code, 4);
        h[3 * i + 2] = make_float4(0.7f, fm, fc, exp2f(8.0f * Cq));          // b, lane mask, row code, h
    }
    float4 *d_entries;
    float *d_out;
    cudaMalloc(&d_entries, sizeof(float4) * kEntries * 3);
    cudaMalloc(&d_out, 16);
    cudaMemcpy(d_entries, h, sizeof(float4) * kEntries * 3, cudaMemcpyHostToDevice);
    const int iters = 200;
    run<0>(d_entries, d_out, iters, sms, ghz, "M0  blend of 4 row pairs (FMUL2, 3 FFMA2, FADD2 each)", 20, 0);
    run<10>(d_entries, d_out, iters, sms, ghz, "M0a blend, T updated by an independent FFMA2 (+4 neg)", 24, 0);
    run<11>(d_entries, d_out, iters, sms, ghz, "M0b blend, blue channel as 2 scalar FFMA per pair", 16, 8);
    run<12>(d_entries, d_out, iters, sms, ghz, "M0c blend, the 4 pairs interleaved operation by operation", 20, 0);
    run<13>(d_entries, d_out, iters, sms, ghz, "M0d blend, colours as register pairs (c, c), not Rn.F32 broadcast", 20, 0);
    run<1>(d_entries, d_out, iters, sms, ghz, "M1  blend + recurrence steps (F *= G, G *= H)", 25, 0);
    run<2>(d_entries, d_out, iters, sms, ghz, "M2  M1 + exponents (3 FFMA2, 3 scalar) + 4 MUFU.EX2", 28, 3);
    run<3>(d_entries, d_out, iters, sms, ghz, "M3  the whole recurrence-path iteration (LDS, prologue, band test)", 28, 8);
    return 0;
}
