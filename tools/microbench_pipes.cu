// Issue-rate model of the FP32 pipe of a B200 scheduler for mixes of packed (f32x2) and scalar
// operations: what the raster kernel's composite loop is made of.  Independent chains, 8 warps per
// scheduler (8 CTAs x 128 threads per SM), no memory traffic.  Output: cycles per pattern per
// scheduler and the implied cost of each operation.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/microbench_pipes.bin tools/microbench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f2_t;

// packed ops on accumulator k (64-bit), scalar ops on accumulator k (32-bit), ALU / XU fillers
#define P_FMA(k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(a2), "l"(b2));
#define P_FMA3(k) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[k]) : "l"(p[(k + 5) % 12]), "l"(a2));
// the blend's accumulate form: 64-bit x, a 32-bit multiplicand broadcast to both halves (SASS Rn.F32)
#define P_FMAB(k, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[k]) : "l"(p[(k + 5) % 12]), "l"(bc[c]));
// the same with the broadcast operand held as a genuine register pair
#define P_FMAQ(k, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[k]) : "l"(p[(k + 5) % 12]), "l"(qc[c]));
#define P_MUL(k) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(a2));
#define P_ADD(k) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(b2));
#define S_FMA(k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[k]) : "f"(a), "f"(b));
#define S_ADD(k) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[k]) : "f"(b));
#define I_ADD(k) asm volatile("add.s32 %0, %0, %1;" : "+r"(i[k]) : "r"(ia));
#define I_LOP(k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(i[k]) : "r"(ia));
#define X_EX2(k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[k]));

#define KERNEL(name, BODY)                                                                          \
    __global__ void __launch_bounds__(128, 8) name(float *out, int iters, f2_t a2, f2_t b2, float a, float b, int ia) \
    {                                                                                               \
        f2_t p[12];                                                                                 \
        float s[8];                                                                                 \
        int i[8];                                                                                   \
        f2_t bc[3], qc[3];                                                                          \
        _Pragma("unroll") for (int k = 0; k < 3; ++k) {                                             \
            const float c = a * (float)(k + 1);                                                     \
            asm("mov.b64 %0, {%1, %1};" : "=l"(bc[k]) : "f"(c));                                    \
            const float lo = c, hi = c + b * 0.0f; /* same value, but not provably: two registers */  \
            asm("mov.b64 %0, {%1, %2};" : "=l"(qc[k]) : "f"(lo), "f"(hi));                          \
        }                                                                                           \
        _Pragma("unroll") for (int k = 0; k < 12; ++k) p[k] = (f2_t)(threadIdx.x + k) * 0x3f8000003f800000ull; \
        _Pragma("unroll") for (int k = 0; k < 8; ++k) s[k] = (float)(threadIdx.x + k), i[k] = threadIdx.x * k; \
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { BODY }                             \
        f2_t x = 0;                                                                                 \
        float t = 0.0f;                                                                             \
        int j = 0;                                                                                  \
        _Pragma("unroll") for (int k = 0; k < 12; ++k) x ^= p[k];                                   \
        _Pragma("unroll") for (int k = 0; k < 8; ++k) t += s[k], j ^= i[k];                         \
        if (x == 0x123456789abcdefull || t == 12345.678f || j == 0x7654321) out[0] = 1.0f;          \
    }

// 12 packed
#define B_P12 P_FMA(0) P_FMA(1) P_FMA(2) P_FMA(3) P_FMA(4) P_FMA(5) P_FMA(6) P_FMA(7) P_FMA(8) P_FMA(9) P_FMA(10) P_FMA(11)
KERNEL(k_p12, B_P12 B_P12)
KERNEL(k_p12_3op, P_FMA3(0) P_FMA3(1) P_FMA3(2) P_FMA3(3) P_FMA3(4) P_FMA3(5) P_FMA3(6) P_FMA3(7) P_FMA3(8) P_FMA3(9) P_FMA3(10) P_FMA3(11)
                  P_FMA3(0) P_FMA3(1) P_FMA3(2) P_FMA3(3) P_FMA3(4) P_FMA3(5) P_FMA3(6) P_FMA3(7) P_FMA3(8) P_FMA3(9) P_FMA3(10) P_FMA3(11))
#define B_FB12 P_FMAB(0, 0) P_FMAB(1, 1) P_FMAB(2, 2) P_FMAB(3, 0) P_FMAB(4, 1) P_FMAB(5, 2) P_FMAB(6, 0) P_FMAB(7, 1) P_FMAB(8, 2) P_FMAB(9, 0) P_FMAB(10, 1) P_FMAB(11, 2)
#define B_FQ12 P_FMAQ(0, 0) P_FMAQ(1, 1) P_FMAQ(2, 2) P_FMAQ(3, 0) P_FMAQ(4, 1) P_FMAQ(5, 2) P_FMAQ(6, 0) P_FMAQ(7, 1) P_FMAQ(8, 2) P_FMAQ(9, 0) P_FMAQ(10, 1) P_FMAQ(11, 2)
KERNEL(k_fmab, B_FB12 B_FB12)
KERNEL(k_fmaq, B_FQ12 B_FQ12)
// one x for three accumulators, as in a blend (the w operand is reused)
#define P_FMAB3(k, x) asm volatile("fma.rn.f32x2 %0, %3, %4, %0;\n\tfma.rn.f32x2 %1, %3, %5, %1;\n\tfma.rn.f32x2 %2, %3, %6, %2;" : "+l"(p[k]), "+l"(p[k + 1]), "+l"(p[k + 2]) : "l"(p[x]), "l"(bc[0]), "l"(bc[1]), "l"(bc[2]));
#define P_FMAQ3(k, x) asm volatile("fma.rn.f32x2 %0, %3, %4, %0;\n\tfma.rn.f32x2 %1, %3, %5, %1;\n\tfma.rn.f32x2 %2, %3, %6, %2;" : "+l"(p[k]), "+l"(p[k + 1]), "+l"(p[k + 2]) : "l"(p[x]), "l"(qc[0]), "l"(qc[1]), "l"(qc[2]));
KERNEL(k_fmab3, P_FMAB3(0, 9) P_FMAB3(3, 10) P_FMAB3(6, 11) P_FMAB3(0, 10) P_FMAB3(3, 11) P_FMAB3(6, 9) P_FMAB3(0, 11) P_FMAB3(3, 9))
KERNEL(k_fmaq3, P_FMAQ3(0, 9) P_FMAQ3(3, 10) P_FMAQ3(6, 11) P_FMAQ3(0, 10) P_FMAQ3(3, 11) P_FMAQ3(6, 9) P_FMAQ3(0, 11) P_FMAQ3(3, 9))
KERNEL(k_pmul12, P_MUL(0) P_MUL(1) P_MUL(2) P_MUL(3) P_MUL(4) P_MUL(5) P_MUL(6) P_MUL(7) P_MUL(8) P_MUL(9) P_MUL(10) P_MUL(11)
                 P_MUL(0) P_MUL(1) P_MUL(2) P_MUL(3) P_MUL(4) P_MUL(5) P_MUL(6) P_MUL(7) P_MUL(8) P_MUL(9) P_MUL(10) P_MUL(11))
KERNEL(k_padd12, P_ADD(0) P_ADD(1) P_ADD(2) P_ADD(3) P_ADD(4) P_ADD(5) P_ADD(6) P_ADD(7) P_ADD(8) P_ADD(9) P_ADD(10) P_ADD(11)
                 P_ADD(0) P_ADD(1) P_ADD(2) P_ADD(3) P_ADD(4) P_ADD(5) P_ADD(6) P_ADD(7) P_ADD(8) P_ADD(9) P_ADD(10) P_ADD(11))
// 24 scalar
#define B_S8 S_FMA(0) S_FMA(1) S_FMA(2) S_FMA(3) S_FMA(4) S_FMA(5) S_FMA(6) S_FMA(7)
KERNEL(k_s24, B_S8 B_S8 B_S8)
// 24 packed + 8 scalar, the scalars in one block / in pairs / one by one
KERNEL(k_p24_s8_block, B_P12 B_P12 B_S8)
KERNEL(k_p24_s8_pairs, P_FMA(0) P_FMA(1) P_FMA(2) P_FMA(3) P_FMA(4) P_FMA(5) S_FMA(0) S_FMA(1) P_FMA(6) P_FMA(7) P_FMA(8) P_FMA(9) P_FMA(10) P_FMA(11) S_FMA(2) S_FMA(3)
                       P_FMA(0) P_FMA(1) P_FMA(2) P_FMA(3) P_FMA(4) P_FMA(5) S_FMA(4) S_FMA(5) P_FMA(6) P_FMA(7) P_FMA(8) P_FMA(9) P_FMA(10) P_FMA(11) S_FMA(6) S_FMA(7))
KERNEL(k_p24_s8_single, P_FMA(0) P_FMA(1) P_FMA(2) S_FMA(0) P_FMA(3) P_FMA(4) P_FMA(5) S_FMA(1) P_FMA(6) P_FMA(7) P_FMA(8) S_FMA(2) P_FMA(9) P_FMA(10) P_FMA(11) S_FMA(3)
                        P_FMA(0) P_FMA(1) P_FMA(2) S_FMA(4) P_FMA(3) P_FMA(4) P_FMA(5) S_FMA(5) P_FMA(6) P_FMA(7) P_FMA(8) S_FMA(6) P_FMA(9) P_FMA(10) P_FMA(11) S_FMA(7))
// 24 packed + 8 integer ALU ops / + 8 logic ops / + 4 MUFU
KERNEL(k_p24_i8, B_P12 I_ADD(0) I_ADD(1) I_ADD(2) I_ADD(3) B_P12 I_ADD(4) I_ADD(5) I_ADD(6) I_ADD(7))
KERNEL(k_p24_l8, B_P12 I_LOP(0) I_LOP(1) I_LOP(2) I_LOP(3) B_P12 I_LOP(4) I_LOP(5) I_LOP(6) I_LOP(7))
KERNEL(k_p24_x4, B_P12 X_EX2(0) X_EX2(1) B_P12 X_EX2(2) X_EX2(3))
KERNEL(k_p24_s8_x4, B_P12 B_S8 X_EX2(0) X_EX2(1) B_P12 X_EX2(2) X_EX2(3))
// the raster loop's mix: 28 packed + 8 scalar + 4 MUFU + 16 ALU-ish (tests, selects, loop control)
KERNEL(k_loop_mix, B_S8 I_ADD(0) I_LOP(1) I_ADD(2) I_LOP(3) P_FMA(0) P_FMA(1) P_FMA(2) X_EX2(0) X_EX2(1) X_EX2(2) X_EX2(3) B_P12 B_P12 P_FMA(3)
                   I_ADD(4) I_LOP(5) I_ADD(6) I_LOP(7) I_ADD(0) I_LOP(1) I_ADD(2) I_LOP(3) I_ADD(4) I_LOP(5) I_ADD(6) I_LOP(7))
KERNEL(k_x8, X_EX2(0) X_EX2(1) X_EX2(2) X_EX2(3) X_EX2(4) X_EX2(5) X_EX2(6) X_EX2(7))

typedef void (*kern_t)(float *, int, f2_t, f2_t, float, float, int);

static double g_ghz;
static int g_sms;

static double run(kern_t k, const char *name, int packed, int scalar, int other, float *d_out)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<g_sms * 8, 128>>>(d_out, 100, 0x3f7fbe773f7fbe77ull, 0x3a83126f3a83126full, 0.999f, 0.001f, 3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<<<g_sms * 8, 128>>>(d_out, iters, 0x3f7fbe773f7fbe77ull, 0x3a83126f3a83126full, 0.999f, 0.001f, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * g_ghz * 1e9 / ((double)iters * 8.0);  // per pattern per scheduler (8 warps each)
    printf("%-58s %3d packed %3d scalar %3d other: %7.2f cycles per pattern", name, packed, scalar, other, cyc);
    printf("  (%d instructions, nominal FMA-pipe cycles %d)\n", packed + scalar + other + 2, 2 * packed + scalar);
    return cyc;
}

int main()
{
    int khz = 0;
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    g_ghz = khz / 1e6;
    float *d_out;
    cudaMalloc(&d_out, 16);
    printf("device: %d SMs at %.3f GHz; 8 warps per scheduler; cycles are per scheduler\n", g_sms, g_ghz);
    const double p = run(k_p12, "FFMA2 d = d*a + b (a, b loop-invariant)", 24, 0, 0, d_out) / 24;
    run(k_p12_3op, "FFMA2 d = x*a + d (x another accumulator)", 24, 0, 0, d_out);
    run(k_fmab, "FFMA2 d = x*c + d, c a 32-bit register broadcast (Rn.F32)", 24, 0, 0, d_out);
    run(k_fmaq, "FFMA2 d = x*c + d, c a register pair holding (c, c)", 24, 0, 0, d_out);
    run(k_fmab3, "3 FFMA2 per x (d_r,g,b += x*c_r,g,b), c broadcast", 24, 0, 0, d_out);
    run(k_fmaq3, "3 FFMA2 per x (d_r,g,b += x*c_r,g,b), c register pairs", 24, 0, 0, d_out);
    run(k_pmul12, "FMUL2", 24, 0, 0, d_out);
    run(k_padd12, "FADD2", 24, 0, 0, d_out);
    const double s = run(k_s24, "FFMA (scalar)", 0, 24, 0, d_out) / 24;
    printf("   -> packed %.3f cycles, scalar %.3f cycles each\n", p, s);
    const double m1 = run(k_p24_s8_block, "24 FFMA2 + 8 FFMA, the scalars in one block", 24, 8, 0, d_out);
    const double m2 = run(k_p24_s8_pairs, "24 FFMA2 + 8 FFMA, the scalars in pairs", 24, 8, 0, d_out);
    const double m3 = run(k_p24_s8_single, "24 FFMA2 + 8 FFMA, the scalars one by one", 24, 8, 0, d_out);
    printf("   -> a scalar FFMA among packed ones costs %.2f / %.2f / %.2f cycles (block / pairs / single)\n",
           (m1 - 24 * p) / 8, (m2 - 24 * p) / 8, (m3 - 24 * p) / 8);
    const double i1 = run(k_p24_i8, "24 FFMA2 + 8 IADD", 24, 0, 8, d_out);
    const double l1 = run(k_p24_l8, "24 FFMA2 + 8 LOP3", 24, 0, 8, d_out);
    printf("   -> an integer add among packed ones costs %.2f cycles, a logic op %.2f\n", (i1 - 24 * p) / 8, (l1 - 24 * p) / 8);
    const double x1 = run(k_p24_x4, "24 FFMA2 + 4 MUFU.EX2", 24, 0, 4, d_out);
    printf("   -> a MUFU.EX2 among packed ones costs %.2f cycles\n", (x1 - 24 * p) / 4);
    run(k_x8, "8 MUFU.EX2", 0, 0, 8, d_out);
    run(k_p24_s8_x4, "24 FFMA2 + 8 FFMA + 4 MUFU.EX2", 24, 8, 4, d_out);
    run(k_loop_mix, "the composite loop's mix: 28 FFMA2 + 8 FFMA + 4 MUFU + 16 ALU", 28, 8, 20, d_out);
    return 0;
}
