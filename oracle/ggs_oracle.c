/*
 * ggs_oracle.c -- CPU restatement of the render + fitness hot path of
 * josedelrey/genetic-gaussian-splats.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and there only as the checker
 * (or as the reported CPU baseline), never as the thing shipped.
 *
 * Parity status: PINNED against the reference's own code run in the build
 * container: tests/golden/make_golden.py imports /root/reference (torch CPU ops
 * + the reference Triton kernel under TRITON_INTERPRET=1) and writes the
 * fixtures under tests/golden/ (npz) that tests/test_oracle_golden.py replays against
 * this file.  The reference itself ships no tests or golden vectors
 * (SURVEY.md section 8c).
 *
 * Every function cites the reference file:line it restates.  All arithmetic is
 * fp32, one rounding per reference torch op; build with -ffp-contract=off so
 * the compiler never fuses a mul+add the reference performs as two ops.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GGS_MODE_PLAIN 0
#define GGS_MODE_MASK 1
#define GGS_MODE_BOOST 2

static inline float clampf(float v, float lo, float hi)
{
    /* torch.clamp: min(max(v, lo), hi); NaN propagates */
    if (v != v)
        return v;
    return v < lo ? lo : (v > hi ? hi : v);
}

static inline float maxf_nan(float v, float lo)
{
    if (v != v)
        return v;
    return v < lo ? lo : v;
}

int ggs_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ggs_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/*
 * Axes-angle genome row -> Cholesky genome row.
 * Reference: modules/encode.py:5-24 (axes_angle_to_cholesky),
 *            modules/encode.py:28-59 (genome_to_renderer),
 *            modules/encode.py:63-79 (genome_to_renderer_batched).
 * in : rows x cols (cols >= 9)  (x, y, log sx, log sy, theta, r, g, b, alpha)
 * out: rows x 9                 (x, y, log l11, log l22, l21, r, g, b, alpha)
 */
static inline void encode_row(const float *g, float *o)
{
    float sx = expf(g[2]);             /* encode.py:6 */
    float sy = expf(g[3]);             /* encode.py:7 */
    float c = cosf(g[4]);              /* encode.py:8 */
    float s = sinf(g[4]);              /* encode.py:9 */
    float sx2 = sx * sx, sy2 = sy * sy; /* sigma**2 == sigma*sigma in torch */
    float c2 = c * c, s2 = s * s;
    float vxx = sx2 * c2 + sy2 * s2;   /* encode.py:12 */
    float vxy = ((sx2 - sy2) * s) * c; /* encode.py:13 */
    float vyy = sx2 * s2 + sy2 * c2;   /* encode.py:14 */
    const float eps = 1e-12f;          /* encode.py:16 */
    float l11 = sqrtf(maxf_nan(vxx, eps));             /* encode.py:17 */
    float l21 = vxy / l11;                             /* encode.py:18 */
    float l22 = sqrtf(maxf_nan(vyy - l21 * l21, eps)); /* encode.py:19 */
    o[0] = g[0];
    o[1] = g[1];
    o[2] = logf(l11); /* encode.py:21 */
    o[3] = logf(l22); /* encode.py:22 */
    o[4] = l21;       /* encode.py:23 */
    for (int k = 5; k < 9; ++k)
        o[k] = clampf(g[k], 0.0f, 255.0f); /* encode.py:57,77 */
}

void ggs_oracle_encode(const float *axes, float *chol, int64_t rows, int cols)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i)
        encode_row(axes + i * cols, chol + i * 9);
}

typedef struct {
    float cx, cy, sxx, sxy, syy, rc, gc, bc, a;
    int x0, x1, y0, y1;
} splat_t;

/*
 * Cholesky genome row -> render record.
 * Reference: modules/render.py:9-47 (_preprocess_genome).
 */
static inline void decode_row(const float *g, int H, int W, float k_sigma, splat_t *s)
{
    float maxx = (float)(W - 1), maxy = (float)(H - 1); /* render.py:14 */
    float cx = clampf(g[0], 0.0f, 1.0f) * maxx;        /* render.py:15 */
    float cy = clampf(g[1], 0.0f, 1.0f) * maxy;        /* render.py:16 */
    float l11 = maxf_nan(expf(g[2]), 1e-6f);           /* render.py:19 */
    float l22 = maxf_nan(expf(g[3]), 1e-6f);           /* render.py:20 */
    float l21 = g[4];                                  /* render.py:21 */
    float hx = maxf_nan(k_sigma * fabsf(l11), 1.0f);   /* render.py:24 */
    float hy = maxf_nan(k_sigma * (fabsf(l21) + fabsf(l22)), 1.0f); /* render.py:25 */
    s->x0 = (int)floorf(clampf(cx - hx, 0.0f, maxx));  /* render.py:27 */
    s->x1 = (int)ceilf(clampf(cx + hx, 0.0f, maxx));   /* render.py:28 */
    s->y0 = (int)floorf(clampf(cy - hy, 0.0f, maxy));  /* render.py:29 */
    s->y1 = (int)ceilf(clampf(cy + hy, 0.0f, maxy));   /* render.py:30 */
    float i11 = (1.0f / l11) * 1.0f;                   /* render.py:32 (reciprocal * 1.0) */
    float i22 = (1.0f / l22) * 1.0f;                   /* render.py:33 */
    float i21 = (-l21) * (i11 * i22);                  /* render.py:34 */
    s->sxx = i11 * i11 + i21 * i21;                    /* render.py:36 */
    s->sxy = i21 * i22;                                /* render.py:37 */
    s->syy = i22 * i22;                                /* render.py:38 */
    s->rc = clampf(g[5], 0.0f, 255.0f) / 255.0f;       /* render.py:40 */
    s->gc = clampf(g[6], 0.0f, 255.0f) / 255.0f;       /* render.py:41 */
    s->bc = clampf(g[7], 0.0f, 255.0f) / 255.0f;       /* render.py:42 */
    s->a = clampf(g[8], 0.0f, 255.0f) / 255.0f;        /* render.py:43 */
    s->cx = cx;
    s->cy = cy;
}

/*
 * Decode rows into the reference's 13 arrays.
 * out_f: [9][rows] = cx, cy, sxx, sxy, syy, rc, gc, bc, a
 * out_i: [4][rows] = x0, x1, y0, y1
 */
void ggs_oracle_decode(const float *chol, int64_t rows, int cols, int H, int W, float k_sigma,
                       float *out_f, int32_t *out_i)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i) {
        splat_t s;
        decode_row(chol + i * cols, H, W, k_sigma, &s);
        out_f[0 * rows + i] = s.cx;
        out_f[1 * rows + i] = s.cy;
        out_f[2 * rows + i] = s.sxx;
        out_f[3 * rows + i] = s.sxy;
        out_f[4 * rows + i] = s.syy;
        out_f[5 * rows + i] = s.rc;
        out_f[6 * rows + i] = s.gc;
        out_f[7 * rows + i] = s.bc;
        out_f[8 * rows + i] = s.a;
        out_i[0 * rows + i] = s.x0;
        out_i[1 * rows + i] = s.x1;
        out_i[2 * rows + i] = s.y0;
        out_i[3 * rows + i] = s.y1;
    }
}

/*
 * Composite one candidate, dense genome-order form of the reference's tile
 * kernel: modules/render.py:157-200 (per-pixel loop, AABB mask :175-177,
 * falloff :189-192, "over" blend :194-196), background fill :236-237,
 * final clamp :252.  Tile binning (:51-118) is pure culling and does not
 * change the result (SURVEY.md finding 3), so it has no counterpart here.
 * img: [H][W][3]
 */
static int64_t composite_one(const float *chol, int N, int cols, int H, int W, float k_sigma,
                             const float *bg, float *img)
{
    int64_t pairs = 0;
    for (int64_t p = 0; p < (int64_t)H * W; ++p) {
        img[3 * p + 0] = bg[0];
        img[3 * p + 1] = bg[1];
        img[3 * p + 2] = bg[2];
    }
    for (int n = 0; n < N; ++n) {
        splat_t s;
        decode_row(chol + (int64_t)n * cols, H, W, k_sigma, &s);
        if (s.x1 < s.x0 || s.y1 < s.y0)
            continue;
        pairs += (int64_t)(s.x1 - s.x0 + 1) * (s.y1 - s.y0 + 1);
        float sxy2 = 2.0f * s.sxy; /* render.py:191 "2.0 * sxy" */
        for (int Y = s.y0; Y <= s.y1; ++Y) {
            float qy = (float)Y - s.cy; /* render.py:190 */
            float qyy = qy * qy;
            float *row = img + ((int64_t)Y * W) * 3;
            for (int X = s.x0; X <= s.x1; ++X) {
                float qx = (float)X - s.cx; /* render.py:189 */
                float quad = s.sxx * (qx * qx) + sxy2 * (qx * qy) + s.syy * qyy; /* :191 */
                float f = expf(-0.5f * quad) * s.a; /* render.py:192 */
                float omf = 1.0f - f;
                float *px = row + 3 * X;
                px[0] = omf * px[0] + f * s.rc; /* render.py:194 */
                px[1] = omf * px[1] + f * s.gc; /* render.py:195 */
                px[2] = omf * px[2] + f * s.bc; /* render.py:196 */
            }
        }
    }
    for (int64_t p = 0; p < (int64_t)H * W * 3; ++p)
        img[p] = clampf(img[p], 0.0f, 1.0f); /* render.py:252 */
    return pairs;
}

/*
 * Batched render of Cholesky-layout genomes.
 * Reference: modules/render.py:204-252 (render_splats_rgb_triton).
 * images: [B][H][W][3].  Returns the number of in-AABB (pixel, splat) pairs,
 * the algorithmic work unit of SURVEY.md section 8d.
 */
int64_t ggs_oracle_render(const float *chol, int B, int N, int cols, int H, int W, float k_sigma,
                          const float *bg, float *images)
{
    int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int b = 0; b < B; ++b)
        total += composite_one(chol + (int64_t)b * N * cols, N, cols, H, W, k_sigma, bg,
                               images + (int64_t)b * H * W * 3);
    return total;
}

/*
 * Fitness of one rendered candidate.
 * Reference: modules/fitness.py:16-31.
 *   plain : mean over (H,W,3) of d^2                                  (:19)
 *   mask  : sum_{h,w,c} d^2 w / (sum_{h,w} w + 1e-12)                  (:29-31)
 *   boost : mean_{h,w,c}(d^2 wb) / (mean_{h,w}(wb) + 1e-12),
 *           wb = 1 + beta*clamp(w,0,1)                                (:23-27)
 * d^2 and the products are fp32 per element as in torch; the reductions are
 * accumulated in double (torch's fp32 tree reduction order is unspecified;
 * parity tolerance on fitness is 1e-5 relative).
 */
static float fitness_one(const float *img, const float *target, const float *mask, int H, int W,
                         int mode, float beta)
{
    double num = 0.0, den = 0.0;
    int64_t P = (int64_t)H * W;
    for (int64_t p = 0; p < P; ++p) {
        float wgt = 1.0f;
        if (mode == GGS_MODE_MASK)
            wgt = mask[p];
        else if (mode == GGS_MODE_BOOST)
            wgt = 1.0f + beta * clampf(mask[p], 0.0f, 1.0f);
        den += (double)wgt;
        for (int c = 0; c < 3; ++c) {
            float d = img[3 * p + c] - target[3 * p + c];
            float d2 = d * d;
            num += (double)(mode == GGS_MODE_PLAIN ? d2 : d2 * wgt);
        }
    }
    if (mode == GGS_MODE_PLAIN)
        return (float)(num / (double)(3 * P));
    if (mode == GGS_MODE_MASK)
        return (float)num / ((float)den + 1e-12f);
    return (float)(num / (double)(3 * P)) / ((float)(den / (double)P) + 1e-12f);
}

/*
 * Batched fitness of axes-angle genomes: stack -> encode -> render -> score.
 * Reference: modules/fitness.py:8-31 (fitness_many).
 * images_out may be NULL.  pairs_out (may be NULL) receives the in-AABB pair
 * count.  Returns 0 on success.
 */
int ggs_oracle_fitness(const float *axes, int B, int N, int cols, int H, int W, float k_sigma,
                       const float *target, const float *mask, int mode, float beta,
                       float *fitness, float *images_out, int64_t *pairs_out)
{
    const float bg[3] = {1.0f, 1.0f, 1.0f}; /* render.py:209 default, never overridden */
    if (mode != GGS_MODE_PLAIN && mask == NULL)
        return -1;
    int64_t total = 0;
    int fail = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total) reduction(| : fail)
    for (int b = 0; b < B; ++b) {
        float *chol = (float *)malloc((size_t)N * 9 * sizeof(float));
        float *img = images_out ? images_out + (int64_t)b * H * W * 3
                                : (float *)malloc((size_t)H * W * 3 * sizeof(float));
        if (!chol || !img) {
            fail |= 1;
        } else {
            for (int n = 0; n < N; ++n)
                encode_row(axes + ((int64_t)b * N + n) * cols, chol + (int64_t)n * 9);
            total += composite_one(chol, N, 9, H, W, k_sigma, bg, img);
            fitness[b] = fitness_one(img, target, mask, H, W, mode, beta);
        }
        free(chol);
        if (!images_out)
            free(img);
    }
    if (pairs_out)
        *pairs_out = total;
    return fail ? -2 : 0;
}

/* Fitness of already-rendered images (used to score golden images). */
void ggs_oracle_score(const float *images, int B, int H, int W, const float *target,
                      const float *mask, int mode, float beta, float *fitness)
{
    for (int b = 0; b < B; ++b)
        fitness[b] = fitness_one(images + (int64_t)b * H * W * 3, target, mask, H, W, mode, beta);
}
