/*
 * ggs_b200.h -- C ABI of the B200-native render + fitness hot path of
 * genetic-gaussian-splats (libggs_b200.so, built from
 * genetic-gaussian-splats_b200/csrc/ for sm_100a).
 *
 * This header is the drop-in boundary: plain pointers and sizes, no torch types.
 * The reference is pure Python and has no FFI of its own; each entry point below
 * names the reference function it replaces (citations into /root/reference).  The
 * Python binding a maintainer adds (ctypes) is shown in INTEGRATION.md and shipped in
 * genetic-gaussian-splats_b200/ggs_b200/native.py.
 *
 * Conventions
 *   - all tensors are dense row-major float32 unless stated otherwise;
 *   - "d_" pointers are device pointers on the current CUDA device, "h_" pointers are
 *     host pointers (pinned memory makes the copies asynchronous, pageable is accepted);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     device-pointer entries only enqueue work: they never synchronise or allocate;
 *   - every function returns GGS_OK (0) or a negative GGS_E* code; ggs_last_error()
 *     returns a thread-local message for the last failure;
 *   - inputs are never written; outputs are fully overwritten.
 *
 * Genome layouts (one row per splat, `cols` >= 9 floats per row, extra columns ignored)
 *   axes-angle : x, y, log sigma_x, log sigma_y, theta, r, g, b, alpha   (population.py:27-43)
 *   Cholesky   : x, y, log l11,     log l22,     l21,   r, g, b, alpha   (encode.py:35-57)
 *   x, y in [0,1] (fraction of W-1, H-1); r, g, b, alpha in [0,255].
 */
#ifndef GGS_B200_H
#define GGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGS_ABI_VERSION 1

#define GGS_OK 0
#define GGS_EINVAL (-1)    /* bad argument (shape, null pointer, mode)          */
#define GGS_ECUDA (-2)     /* a CUDA runtime call or kernel launch failed       */
#define GGS_EWORKSPACE (-3) /* workspace too small; see ggs_workspace_bytes()    */
#define GGS_ENODEVICE (-4) /* no sm_100 device visible                          */

#define GGS_LAYOUT_AXES_ANGLE 0
#define GGS_LAYOUT_CHOLESKY 1

/* fitness.py:18-31 */
#define GGS_MODE_PLAIN 0 /* weight_mask is None: mean over (H,W,3) of d^2                    */
#define GGS_MODE_MASK 1  /* sum(d^2 * w) / (sum_{H,W} w + 1e-12)   (denominator not x3)       */
#define GGS_MODE_BOOST 2 /* boost_only: mean(d^2 * wb) / (mean(wb) + 1e-12), wb=1+beta*clamp(w)*/

/* Largest image side the packed int16 AABB supports. */
#define GGS_MAX_SIDE 32768

int ggs_abi_version(void);
const char *ggs_last_error(void);

/* Number of CUDA devices visible to the library, or a negative error. */
int ggs_device_count(void);

/*
 * Device scratch needed by ggs_render / ggs_fitness for a batch of B candidates of N
 * splats on an H x W image: decoded splat records, packed AABBs, per-tile partial sums
 * and per-candidate completion counters.  The caller owns the buffer (e.g. a torch
 * uint8 tensor) and may reuse it across calls on the same stream.  It need not be
 * initialised.
 */
size_t ggs_workspace_bytes(int B, int N, int H, int W);

/*
 * genome_to_renderer_batched (modules/encode.py:63-79, via :28-59 and :5-24).
 * d_axes: rows x cols (cols >= 9)  ->  d_chol: rows x 9.  Colours/alpha clamped to
 * [0,255].  Same fp32 operation order as the reference (no FMA contraction).
 */
int ggs_encode(const float *d_axes, int64_t rows, int cols, float *d_chol, void *stream);

/*
 * _preprocess_genome (modules/render.py:9-47), for tests and tools.
 * d_genomes: rows x cols in `layout`; outputs in the reference's naming:
 *   d_out_f: [9][rows] = cx, cy, sxx, sxy, syy, rc, gc, bc, a
 *   d_out_i: [4][rows] = x0, x1, y0, y1   (int32, inclusive AABB)
 */
int ggs_decode(const float *d_genomes, int layout, int64_t rows, int cols, int H, int W,
               float k_sigma, float *d_out_f, int32_t *d_out_i, void *stream);

/*
 * render_splats_rgb_triton (modules/render.py:204-252).
 * d_genomes: [B][N][cols] in `layout` (the reference entry takes Cholesky layout);
 * h_background: 3 floats on the host (reference default 1,1,1);
 * d_images: [B][H][W][3] float32, clamped to [0,1].
 */
int ggs_render(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
               float k_sigma, const float *h_background, float *d_images, void *d_workspace,
               size_t workspace_bytes, void *stream);

/*
 * Same render, 8-bit output for frames and final images: d_images_u8 [B][H][W][3] =
 * (uint8)(image * 255), truncating like the reference's
 * (img.clamp(0,1).cpu().numpy() * 255).astype("uint8")  (modules/utils.py:57, run_ggs.py:71);
 * a quarter of the bytes cross PCIe.
 */
int ggs_render_u8(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
                  float k_sigma, const float *h_background, unsigned char *d_images_u8,
                  void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * fitness_many (modules/fitness.py:8-31): encode + decode + render + masked squared
 * error fused; candidate images never touch HBM unless d_images is given.
 * d_genomes: [B][N][cols] in `layout` (the reference passes axes-angle);
 * d_target: [H][W][3] in [0,1]; d_mask: [H][W] or NULL (required unless GGS_MODE_PLAIN);
 * d_fitness: [B] (lower is better); d_images: [B][H][W][3] or NULL.
 * Background is white as in the reference (render.py:209).  The per-candidate
 * reduction order is fixed, so results are bit-reproducible run to run and, for a given
 * `split` (see ggs_fitness_ex below), independent of how a population is split into calls.
 */
int ggs_fitness(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
                float k_sigma, const float *d_target, const float *d_mask, int mode,
                float boost_beta, float *d_fitness, float *d_images, void *d_workspace,
                size_t workspace_bytes, void *stream);

/*
 * Small batches (simulated annealing's neighbours, a 32-individual GA, single frames).  One CTA
 * per (candidate, 32x32 tile) leaves most of a B200 idle when B * tiles is a few hundred, so the
 * library can give a tile to a thread-block cluster of `split` = 2, 4 or 8 CTAs: CTA k composites
 * the k-th segment of the genome and the partial (colour, transmittance) states are folded in
 * genome order through distributed shared memory ("over" is associative).  (A variant that also
 * fuses the decode into that launch exists behind the "fuse" option; it measured slower and is off.)
 * Every entry point picks `split` from B (ggs_choose_split, from measurements: 8 while the grid
 * stays within half a wave, else 2 or 4 up to about 7 CTAs per SM while the unsplit grid has
 * fewer than 3 per SM, segments of at least 16 splats, and 1 for genomes of more than 1,536 splats, whose
 * segments would lose the saturation stop).  The fold changes the floating-point association, so results for different
 * `split` agree to ~1e-7 but not bit for bit: a caller that evaluates ONE population in several
 * calls or on several GPUs and wants the bits of a single call passes the split of the whole
 * population explicitly (ggs_fitness_ex; ggs_ctx_fitness_host and the engines do).
 */
int ggs_choose_split(int B, int N, int H, int W);
/* ggs_fitness with an explicit split: 0 = ggs_choose_split(B, ...), else 1, 2, 4 or 8. */
int ggs_fitness_ex(const float *d_genomes, int layout, int B, int N, int cols, int H, int W,
                   float k_sigma, const float *d_target, const float *d_mask, int mode,
                   float boost_beta, float *d_fitness, float *d_images, void *d_workspace,
                   size_t workspace_bytes, int split, void *stream);
/*
 * Process-wide switches, for A/B timing and tests; the defaults come from the environment
 * variables read at first use.  "pdl" (GGS_B200_PDL, default 1): programmatic dependent launch
 * between the kernels of a step.  "split" (GGS_B200_SPLIT, default 0 = automatic): force 1, 2, 4
 * or 8 wherever the caller does not pass one.  "fuse" (GGS_B200_FUSE, default 0 = never): 1 = decode
 * inside the raster launch whenever a segment fits the list, -1 = only for single-wave grids.
 * "tile_order" (GGS_B200_TILE_ORDER, default 1): grids between two CTAs per SM and sixteen waves
 * launch the image's tiles from the centre outwards, tile-major (the CTAs an SM is dealt last are
 * then the cheap ones: a border tile lists ~60 % of the splats of an interior one); 0 =
 * candidate-major always.  Results are bit-identical either way.
 */
int ggs_set_option(const char *name, int value);
/* The centre-out tile order itself, for inspection and tests (host only, no GPU work): writes
 * (tx, ty) of the tile of rank r to out_xy[2r], out_xy[2r + 1] for r = 0 .. ntx * nty - 1. */
int ggs_tile_order(int ntx, int nty, int *out_xy);

/* ---- host-buffer path (fitness_population, modules/fitness.py:35-48) ------------- */

typedef struct ggs_ctx ggs_ctx;

/* Creates a context on `device`: two streams, grow-only device buffers. */
int ggs_ctx_create(int device, ggs_ctx **out);
void ggs_ctx_destroy(ggs_ctx *ctx);

/*
 * One-time upload of the target [H][W][3] and optional weight mask [H][W]; both stay
 * resident on the device for all later ggs_ctx_fitness_host calls.
 */
int ggs_ctx_set_target(ggs_ctx *ctx, const float *h_target, const float *h_mask, int H, int W);

/*
 * fitness_population (modules/fitness.py:35-48) on host buffers: copies the genomes
 * host->device in slices that overlap with the kernels of the previous slice, runs the
 * fused evaluation, copies the B fitness values back and returns when h_fitness is
 * valid (the counterpart of the reference's `.cpu().tolist()`).
 * h_genomes: [B][N][cols] in `layout`; mode/boost_beta as in ggs_fitness (the mask given
 * to ggs_ctx_set_target is used).
 */
int ggs_ctx_fitness_host(ggs_ctx *ctx, const float *h_genomes, int layout, int B, int N,
                         int cols, float k_sigma, int mode, float boost_beta, float *h_fitness);

/* ---- GA breeding step (callers of the hot path; SURVEY.md section 8f "next" #1) ---- */

/*
 * One generation of selection + crossover + mutation for the whole population in one launch:
 * tournament_selection (modules/genetic.py:8-14) with the shuffle/pairing of
 * algorithm.py:87-100, crossover_uniform (genetic.py:17-21), mutate_individual
 * (genetic.py:32-92, incl. the "at least one gene per group" rule and the size-ordered splat
 * swap) and clamp_genome (utils.py:36-45).
 * d_population: [P][N][cols] axes-angle genomes; d_fitness: [P] (lower is better);
 * d_offspring: [P][N][9], must not alias the population (elitism is the caller's row copy).
 * h_sigma6: annealed mutation sigmas on the host, in the order xy, alog, blog, theta, rgb,
 * alpha (build_mut_sigma, utils.py:31-33); log_scale_lo/hi: log of the legal sigma range.
 * Counter-based RNG: the result is a pure function of (seed, generation, inputs).  Same
 * distributions as the reference, different random streams.
 */
int ggs_ga_breed(const float *d_population, const float *d_fitness, int P, int N, int cols,
                 float *d_offspring, int tour_k, float cxpb, float mutpb, const float *h_sigma6,
                 float log_scale_lo, float log_scale_hi, uint64_t seed, uint32_t generation,
                 void *stream);

/* ---- GA engine: generations without the host in the loop (section 8f "next" #1) ---- */

/*
 * The whole generation of algorithm.py:87-160 on the device: breeding (ggs_ga_breed's kernel),
 * evaluation of the children (the fused render + fitness path), elitism (the n_elite best
 * survive unchanged at the front of the next population, algorithm.py:128-141), the stable
 * ranking of the new fitness vector, the (best so far, mean, median) curve point and the
 * best-individual bookkeeping (algorithm.py:143-160).  ggs_ga_run only ENQUEUES work (four
 * launches per generation on `stream`); nothing is copied to the host until ggs_ga_state.
 * The engine owns its device memory (two generation buffers, target, mask, workspace, curves).
 * Limits: P <= 16384 (the ranking sorts in one CTA's shared memory).
 */
typedef struct ggs_ga ggs_ga;
int ggs_ga_create(int device, int P, int N, int H, int W, int n_elite, int max_generations,
                  ggs_ga **out);
void ggs_ga_destroy(ggs_ga *ga);
/* d_target [H][W][3], d_mask [H][W] or NULL (device pointers, copied); mode / boost_beta /
 * k_sigma as in ggs_fitness. */
int ggs_ga_set_target(ggs_ga *ga, const float *d_target, const float *d_mask, int mode,
                      float boost_beta, float k_sigma, void *stream);
/* Generation 0: copies the population [P][N][cols] (axes-angle), evaluates and ranks it.
 * `seed` keys the counter-based random streams of every later generation. */
int ggs_ga_start(ggs_ga *ga, const float *d_population, int cols, uint64_t seed, void *stream);
/* Enqueue `count` more generations.  h_sigma6: [count][6] annealed mutation sigmas, one row per
 * generation in ggs_ga_breed's order (the schedule stays with the caller, utils.py:19-33);
 * the other parameters as in ggs_ga_breed.  Does not synchronise. */
int ggs_ga_run(ggs_ga *ga, int count, const float *h_sigma6, int tour_k, float cxpb, float mutpb,
               float log_scale_lo, float log_scale_hi, void *stream);
/* Synchronises `stream` and reports: generations completed, best fitness so far, generations
 * since the last improvement, curve points [curves_from, generation] as (best, mean, median)
 * triples, the best individual [N][9].  Any output pointer may be NULL. */
int ggs_ga_state(ggs_ga *ga, void *stream, int *h_generation, double *h_best_fitness,
                 int *h_no_improve, double *h_curves3, int curves_from, float *h_best_individual);
/* Device pointers to the current population [P][N][9] and its fitness [P]; valid until the
 * next ggs_ga_run. */
int ggs_ga_population(ggs_ga *ga, const float **d_population, const float **d_fitness);

/* ---- simulated annealing on the device ---------------------- */

/*
 * The iteration loop of modules/annealing.py:112-150 on the device, in two schemes
 * (ggs_sa_set_mode):
 *   sequential (default, the reference's chain, annealing.py:121-146): each of the `tries` of an
 *     iteration mutates the state the previous try left behind (the breeding kernel with a
 *     one-individual population and no crossover), is evaluated on its own (B = 1, the raster's
 *     split path) and passes the Metropolis test -- accept when dE <= 0 or u < exp(-dE / T), the
 *     best-so-far test follows every try -- before the next one is proposed: 4 launches per try;
 *   batched (BASELINE config 2, "batched neighbour proposals"): `tries` independently mutated
 *     copies of the current state, ONE evaluation of all of them, the Metropolis tests applied to
 *     them in order: 4 launches per iteration, but a later try no longer starts from an accepted
 *     earlier try of the same iteration, so iterations are not 1:1 with the reference's.
 * ggs_sa_run only ENQUEUES work on `stream`; the temperature schedule, the annealed sigmas and
 * the U[0,1) draws stay with the caller and are passed per iteration.  The engine owns its
 * device memory.  Limits: tries <= 64.
 */
typedef struct ggs_sa ggs_sa;
int ggs_sa_create(int device, int N, int H, int W, int tries, int max_iterations, ggs_sa **out);
void ggs_sa_destroy(ggs_sa *sa);
/* batched_neighbours = 0: the reference's sequential tries (default); != 0: one evaluation of all
 * tries per iteration.  May be changed between ggs_sa_run calls. */
int ggs_sa_set_mode(ggs_sa *sa, int batched_neighbours);
/* As ggs_ga_set_target. */
int ggs_sa_set_target(ggs_sa *sa, const float *d_target, const float *d_mask, int mode,
                      float boost_beta, float k_sigma, void *stream);
/* Iteration 0: copies the state [N][cols] (axes-angle), evaluates it; it is the current and the
 * best state (annealing.py:99-102).  `seed` keys the random streams of the proposals. */
int ggs_sa_start(ggs_sa *sa, const float *d_state, int cols, uint64_t seed, void *stream);
/* Enqueue `count` more iterations.  h_sigma6: [count][6] as in ggs_ga_run; h_temperature:
 * [count]; h_uniform: [count][tries], the draw try k of that iteration compares with
 * exp(-dE / T) when it is uphill.  Does not synchronise. */
int ggs_sa_run(ggs_sa *sa, int count, const float *h_sigma6, const double *h_temperature,
               const double *h_uniform, float mutpb, float log_scale_lo, float log_scale_hi,
               void *stream);
/* Synchronises `stream` and reports: iterations completed, best and current energy, curve
 * points [curves_from, iteration] as (best, current) pairs, the best and the current state
 * [N][9].  Any output pointer may be NULL. */
int ggs_sa_state(ggs_sa *sa, void *stream, int *h_iteration, double *h_best_energy,
                 double *h_current_energy, double *h_curves2, int curves_from, float *h_best_state,
                 float *h_current_state);

/* ---- the population sharded over the GPUs of one box ----------------- */

/*
 * The fitness all-gather without a collective launch.  One process per GPU; every rank creates a
 * ggs_peers on its device, exports a CUDA IPC handle of its receive buffer, the handles are
 * exchanged by the host (torch.distributed, MPI, a file: 64 bytes per rank) and mapped with
 * ggs_peers_connect.  After that the raster kernel itself delivers: the CTA that finishes a
 * candidate stores its fitness into EVERY rank's gathered vector over NVLink, and the launch's
 * last candidate raises a flag on every rank (system-scope release); consumers wait for `world`
 * flags.  All ranks must issue the same sequence of gathers (epochs advance in lockstep); the
 * gathered vector of a call stays valid until the call after the next one.
 * capacity: largest population (floats in the gathered vector); world <= 8.
 */
typedef struct ggs_peers ggs_peers;
#define GGS_IPC_HANDLE_BYTES 64
int ggs_peers_create(int device, int rank, int world, int capacity, ggs_peers **out);
void ggs_peers_destroy(ggs_peers *peers);
int ggs_peers_export(ggs_peers *peers, void *h_handle /* GGS_IPC_HANDLE_BYTES */);
int ggs_peers_connect(ggs_peers *peers, const void *h_handles /* world x GGS_IPC_HANDLE_BYTES, by rank */);
/* Ranks that live in ONE process (one thread or process driving several GPUs, or several groups
 * on one GPU): connect through the objects themselves instead of IPC handles.  all[r] = rank r. */
int ggs_peers_connect_local(ggs_peers *peers, ggs_peers *const *all);
/* Synchronises `stream`; GGS_ECUDA if a wait for a peer's values timed out (a rank died). */
int ggs_peers_status(ggs_peers *peers, void *stream);
/*
 * fitness_many of this rank's shard -- candidates [offset, offset + B) of a population of `total`
 * -- with the result delivered to every rank: enqueues the evaluation (kernel configuration of
 * the WHOLE population, so the gathered vector has the bits of a single-GPU evaluation) and a
 * one-warp wait for the other ranks' values.  *d_gathered: this rank's copy of the whole
 * vector [total], valid in stream order after this call.  B may be 0 (an empty shard).
 */
int ggs_fitness_allgather(ggs_peers *peers, const float *d_genomes, int layout, int B, int N,
                          int cols, int H, int W, float k_sigma, const float *d_target,
                          const float *d_mask, int mode, float boost_beta, int offset, int total,
                          void *d_workspace, size_t workspace_bytes, const float **d_gathered,
                          void *stream);
/*
 * Shards the evaluation of a GA engine over the ranks of `peers` (or NULL: back to one GPU).
 * Every rank runs the same engine calls on the same starting population and seed: breeding,
 * elitism and ranking are replicated (deterministic counter-based streams), rank r evaluates its
 * contiguous slice of the children and the fitness values travel as above -- the select kernel
 * waits for the flags itself, so a generation is still four launches and no host sync.  Call
 * before ggs_ga_start.  The result is bit-identical to the single-GPU engine.
 */
int ggs_ga_set_peers(ggs_ga *ga, ggs_peers *peers);

/* ---- importance mask, the weight_mask input of the fitness ---- */

/*
 * compute_importance_mask (modules/mask.py:29-83) on the device: bilinear resize of the source
 * image to the work size (mask.py:47), Rec.709 luma (mask.py:6-10), multi-scale Sobel energy
 * (mask.py:50-59), 9x9 local variance (mask.py:21-25, 62), the 2nd..98th percentile
 * normalisation with torch.quantile's linear interpolation (mask.py:65-69), mix (mask.py:73),
 * optional box smoothing (mask.py:75-77), gamma / floor / strength (mask.py:79-86).
 * d_image: [H0][W0][3]; image_is_0_255 != 0 divides by 255 first (the reference decides this
 * with `x.max() > 1.5`, mask.py:45; the caller makes that test).  h_edge_scales: n_scales
 * integers >= 1, each <= min(H, W).  smooth: 0 (off) or an odd box size.  The scalar
 * parameters are doubles because the reference does its scalar arithmetic (1.0 - floor,
 * 1.0 - strength) in Python floats before rounding to float32.
 * d_mask: [H][W].  Workspace: ggs_mask_workspace_bytes(H, W), caller-owned like the others.
 */
size_t ggs_mask_workspace_bytes(int H, int W);
int ggs_importance_mask(const float *d_image, int H0, int W0, int H, int W, int image_is_0_255,
                        const int *h_edge_scales, int n_scales, double w_edge, double w_var,
                        double gamma, double floor, int smooth, double strength, float *d_mask,
                        void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- hardware probes used by bench.py for the roofline denominators -------------- */

/*
 * Measures on the current device, with CUDA events: out[0] = FFMA TFLOP/s (scalar fp32
 * FMA, 2 flop each), out[1] = FFMA2 TFLOP/s (packed fma.rn.f32x2), out[2] = MUFU.EX2
 * Gop/s, out[3] = SM count, out[4] = SM clock (MHz) reported by the device attributes.
 */
int ggs_probe_peaks(float *h_out5);

/*
 * Per-kernel timing for bench.py's roofline.  While enabled, every evaluation
 * (ggs_render / ggs_fitness / ggs_ctx_fitness_host) records CUDA events on the caller's
 * stream around its decode launch and around its raster launch (at most 4096 evaluations
 * are kept).  ggs_timing_read waits for the recorded events, returns the summed decode and
 * raster kernel times in milliseconds and the number of evaluations, and clears the log.
 * Not thread-safe; meant for a single benchmarking thread.
 */
int ggs_timing_enable(int enable);

/*
 * Work counters for bench.py's roofline.  While d_counters2 (2 x uint64 on the device, zeroed
 * by the caller) is set, evaluations run an instrumented copy of the raster kernel that adds
 * the number of 2-row x 32-column pixel blocks it blended on the recurrence path to
 * d_counters2[0] and on the exact path to d_counters2[1] (x 64 = pixel-splat pairs actually
 * evaluated, as opposed to the algorithmic in-AABB pairs).  Pass NULL to switch back to the
 * production kernel.  Not thread-safe.
 */
int ggs_stats_target(unsigned long long *d_counters2);
int ggs_timing_read(float *h_decode_ms, float *h_raster_ms, int *h_evaluations);

#ifdef __cplusplus
}
#endif
#endif /* GGS_B200_H */
