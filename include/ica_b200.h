/*
 * ica_b200.h -- C-ABI of the B200-native inverse compositional registration library
 * (libica_b200.so, hand-written sm_100a CUDA; no torch types, plain pointers and sizes).
 *
 * The reference (mfournigault/inverse_compositional_algorithm) has NO FFI/plugin boundary:
 * its public surface is the Python function API of src/ (SURVEY.md 8b).  The entry points
 * below are what a ctypes binding of that API binds; each one cites the reference function
 * it replaces.  The reference-side binding is shown in INTEGRATION.md and implemented in
 * inverse_compositional_algorithm_b200/_native.py.
 *
 * Conventions
 *  - every function returns 0 on success or a negative ica_status; ica_last_error() gives text
 *    (no exceptions cross the ABI);
 *  - images are float32 (or uint8 on the host entry), channels-last [B][H][W][C], C in {1,3};
 *    parameters, errors and all accumulations are float64;
 *  - integer codes equal the reference's Enum values: transform 1..5 = TRANSLATION, EUCLIDEAN,
 *    SIMILARITY, AFFINITY, HOMOGRAPHY (src/transformation.py:8-13); robust 0..4 = QUADRATIC,
 *    TRUNCATED_QUADRATIC, GERMAN_MCCLURE, LORENTZIAN, CHARBONNIER (src/image_optimisation.py:10-15);
 *  - parameter vectors are stored 8 doubles per pair, the first nparams() used, rest zero;
 *  - all device work is enqueued on the caller's stream (void* = cudaStream_t, NULL = default);
 *    nothing synchronises with the host except where stated;
 *  - a plan is not thread-safe; different plans may be used from different threads.
 */
#ifndef ICA_B200_H
#define ICA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ICA_API __attribute__((visibility("default")))
#else
#define ICA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ICA_MAX_PARAMS 8
#define ICA_MAX_SCALES 12
#define ICA_TRAJ_STRIDE 12 /* scale, iter, |dp|, lambda, p[8] */

typedef enum {
  ICA_OK = 0,
  ICA_ERR_INVALID = -1,   /* bad argument (the reference raises ValueError) */
  ICA_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  ICA_ERR_NO_DEVICE = -3, /* no CUDA device visible: there is no CPU fallback */
  ICA_ERR_ALLOC = -4
} ica_status;

/* flags for ica_config.flags */
#define ICA_FLAG_RECORD_TRAJECTORY 1u /* keep (scale, iter, |dp|, lambda, p) per iteration */
#define ICA_FLAG_WRITE_DI_IW 2u       /* produce the DI / Iw images the reference returns */
/* ica_plan_run_host dtype modifier: the host images are RGB [B][H][W][3] and the (one-channel) plan registers their
   luminance Y = 0.2125 R + 0.7154 G + 0.0721 B (skimage.color.rgb2gray weights), converted on the device */
#define ICA_DTYPE_RGB_TO_LUMA 0x10
/* ica_plan_run_host dtype modifier: DI_out / Iw_out are float64 arrays (what the reference returns); the widening from the
   device's float32 happens on the device, before the device->host copy */
#define ICA_DTYPE_OUT_F64 0x20
/* IPOL-faithful options the reference carries but does not use on its default path; with both set the
   quadratic runs reproduce the IPOL C++ console logs stored in the reference's docs/Algortihm Report.md:38-339 */
#define ICA_FLAG_IPOL_PYRAMID 16u     /* levels by zoom.zoom_out (src/zoom.py:29-60) instead of skimage rescale */
#define ICA_FLAG_IPOL_WARP 32u        /* warp domain of bicubic_interpolation_image (src/bicubic_interpolation.py:121-152):
                                         valid iff the projected point lies in [delta, n-1-delta]; no clip; needs delta >= 2 */
#define ICA_FLAG_HOST_LOOP 8u         /* drive the iteration loop from the host (polling) instead of the
                                         default CUDA-graph while node whose condition is set on the device */

typedef struct ica_config {
  int32_t batch;          /* B independent image pairs per run */
  int32_t height, width;  /* level-0 shape (ny, nx) */
  int32_t channels;       /* 1 or 3 (channels-last) */
  int32_t gray_as_rgb;    /* C==1 only: behave as the reference on the gray image replicated x3
                             (it rejects non-RGB input, ica.py:48-49, 300-301) */
  int32_t nscales;        /* pyramid levels, 1 = single scale (ica.py:264-374) */
  double nu;              /* downsampling factor, 0 < nu < 1 (ica.py:333) */
  int32_t transform_type; /* default for every pair; override with ica_plan_set_transform_types */
  int32_t robust_type;    /* 0 with robust_loop==0 -> quadratic loop ica.py:17-133 */
  int32_t robust_loop;    /* 1: per-iteration weighted Hessian (ica.py:135-261) even for QUADRATIC */
  double lambda_;         /* > 0 fixed; <= 0 -> schedule 80 * 0.9^k floored at 5 (ica.py:223,235-238) */
  double tol;             /* stop when |dp| <= tol; must be < 0.01 (ica.py:59-60) */
  int32_t max_iter;       /* constants.MAX_ITER = 30 */
  int32_t delta;          /* width of the discarded frame (ica.py:85-93) */
  int32_t nanifoutside;   /* frame is applied iff nanifoutside != 0 and delta > 0 */
  uint32_t flags;
  int32_t blocks_per_pair; /* 0 = choose from batch size and SM count */
} ica_config;

typedef struct ica_plan ica_plan;

/* ---- library / device --------------------------------------------------------------------- */
ICA_API const char* ica_last_error(void);
ICA_API int ica_version(void);
ICA_API int ica_device_count(void);       /* 0 when no GPU is visible */
ICA_API int ica_set_device(int device);
ICA_API int ica_get_device(int* device_out);  /* the calling thread's current CUDA device */
/* compile-time constants of the native code, for the constants.py parity test:
   out[0..4] = MAX_ITER, LAMBDA_0, LAMBDA_N, LAMBDA_RATIO, prefilter pad (12) */
ICA_API int ica_get_constants(double* out5);

/* ---- plan: pyramid + state + partial buffers for B pairs of one shape ---------------------- */
/* Replaces the allocation/validation preamble of pyramidal_inverse_compositional_algorithm
   (src/inverse_compositional_algorithm.py:300-337). */
ICA_API int ica_plan_create(const ica_config* cfg, ica_plan** plan_out);
ICA_API int ica_plan_destroy(ica_plan* plan);
/* per-pair transform types (e.g. the similarity/affinity mix of BASELINE config 3) */
ICA_API int ica_plan_set_transform_types(ica_plan* plan, const int32_t* types, int32_t count);
/* level shapes chosen by zoom.zoom_size (src/zoom.py:8-22): nx[s], ny[s] for s < nscales */
ICA_API int ica_plan_level_shapes(const ica_plan* plan, int32_t* nx_out, int32_t* ny_out);
ICA_API size_t ica_plan_device_bytes(const ica_plan* plan);

/* Whole registration of B pairs, inputs already on the device.  Replaces
   pyramidal_inverse_compositional_algorithm (ica.py:264-374), or with nscales == 1
   inverse_compositional_algorithm (ica.py:17-133) / robust_inverse_compositional_algorithm
   (ica.py:135-261).  I1/I2: float32 [B][H][W][C]; p_inout: double [B][8] initial parameters
   (only used when nscales == 1, like the reference, ica.py:327,372) and final result.
   Asynchronous on `stream` except for the convergence polls (none with ICA_FLAG_GRAPH_LOOP). */
ICA_API int ica_plan_run_device(ica_plan* plan, const float* I1_dev, const float* I2_dev,
                        double* p_inout_dev, void* stream);

/* ---- Row-sharded mode: ONE large pair split across ranks by bands of rows (BASELINE.json configs[4],
   SURVEY.md 8e).  The sums over pixels of ica.py:95-99 (H) and :101,240-246 (b) are the only global
   operation of an iteration; they are linear in the pixels, so each rank gathers the moment sums of its
   band and the caller adds them over ranks (NCCL allreduce of ica_moment_stride() doubles per pair).
   Every rank holds both full images (pyramid and warps need neighbours across the band edge), performs
   the identical solve on the identical reduced moments and therefore keeps identical parameters.
     ica_plan_set_row_shard(plan, rank, nranks)
     ica_plan_shard_begin(plan, I1, I2, p_in, stream)          pyramids, state, first work list
     loop: ica_plan_shard_partial(plan, moments, stream)       K2 on this rank's band -> moments[B][stride]
           <allreduce(moments, SUM) by the caller on the same stream>
           ica_plan_shard_solve(plan, moments, &n_active, stream)   K3; n_active == 0 ends the loop
     ica_plan_shard_finish(plan, p_out, stream)                results (and DI/Iw with the flag)          */
ICA_API int ica_moment_stride(void);
/* band of tile rows [ty0, ty1) that rank owns at a level with tiles_y rows of tiles (no device needed) */
ICA_API int ica_row_band(int32_t tiles_y, int32_t rank, int32_t nranks, int32_t* ty0, int32_t* ty1);
ICA_API int ica_plan_set_row_shard(ica_plan* plan, int32_t rank, int32_t nranks);
ICA_API int ica_plan_shard_begin(ica_plan* plan, const float* I1_dev, const float* I2_dev,
                                 const double* p_in_dev, void* stream);
ICA_API int ica_plan_shard_partial(ica_plan* plan, double* moments_dev, void* stream);
/* n_active_out may be NULL (no synchronisation); otherwise the stream is synchronised */
ICA_API int ica_plan_shard_solve(ica_plan* plan, const double* moments_dev, int32_t* n_active_out, void* stream);
ICA_API int ica_plan_shard_finish(ica_plan* plan, double* p_out_dev, void* stream);

/* Row-sharded mode with the exchange INSIDE the device-side loop (no NCCL call, no host round trip per iteration):
   every rank owns an exchange buffer that the other ranks map through CUDA IPC (NVLink peer access); the solve kernel
   stores the band's 105 moment sums into every rank's buffer, raises a sequence flag, waits for the other ranks' flags
   and adds the sums in rank order.  One process per GPU.
     ica_plan_xchg_create(plan, world, rank, handle64)   allocate + export this rank's buffer (64-byte IPC handle)
     <all-gather the handles among the ranks>
     ica_plan_xchg_connect(plan, handles[world][64])     map the peers
     ica_plan_run_row_sharded(plan, I1, I2, p, stream)   whole registration, one graph launch; identical p on all ranks
     ica_plan_xchg_stats(plan, &mean_us, &count, &error) exchange latency (publish -> all ranks seen); error != 0: a peer
                                                         did not answer within 4 s                                       */
ICA_API int ica_plan_xchg_create(ica_plan* plan, int32_t world, int32_t rank, void* ipc_handle_out64);
ICA_API int ica_plan_xchg_connect(ica_plan* plan, const void* ipc_handles);
ICA_API int ica_plan_run_row_sharded(ica_plan* plan, const float* I1_dev, const float* I2_dev, double* p_inout_dev, void* stream);
ICA_API int ica_plan_xchg_stats(ica_plan* plan, double* mean_us_out, int64_t* count_out, int32_t* error_out);

/* Same, from HOST buffers (what the Python drop-in calls): copies inputs host->device, runs,
   copies results back, synchronises.  dtype: 0 = float32, 1 = uint8, 2 = float64 (converted
   to float32 on the device).  Optional outputs may be NULL.
   err_out[B] = last |dp| (the reference's `error`), iters_out[B][nscales] iterations per scale,
   DI_out/Iw_out float32 [B][H][W][C] as returned by the reference (from the last iteration's
   warp, i.e. before the final update, ica.py:227-251,261). */
ICA_API int ica_plan_run_host(ica_plan* plan, const void* I1_host, const void* I2_host, int32_t dtype,
                      double* p_inout_host, double* err_out, int32_t* iters_out,
                      float* DI_out, float* Iw_out);

/* Device time (CUDA events on the plan's stream) of the last ica_plan_run_host call, from the
   first host->device copy to the last device->host copy */
ICA_API int ica_plan_last_host_run_ms(ica_plan* plan, float* ms_out);

/* Profiling hook: per-CTA %globaltimer stamps (16 per CTA) of the LAST iterate launch's first work
   item; enable != 0 allocates the buffer, host_out (grid*16 int64) receives it; returns the grid size */
ICA_API int ica_plan_debug_timeline(ica_plan* plan, long long* host_out, int32_t enable);

/* Results of the last run (device -> host copies; synchronises the plan's stream). */
ICA_API int ica_plan_get_results(ica_plan* plan, double* p_out, double* err_out, int32_t* iters_out);
/* trajectory of the last run: traj_out[B][nscales*max_iter][ICA_TRAJ_STRIDE], count_out[B] */
ICA_API int ica_plan_get_trajectory(ica_plan* plan, double* traj_out, int32_t* count_out);
/* DI / Iw of the last run as device pointers owned by the plan (ICA_FLAG_WRITE_DI_IW) */
ICA_API int ica_plan_get_di_iw_device(ica_plan* plan, const float** DI_dev, const float** Iw_dev);
/* pyramid level s of image `which` (0 = I1, 1 = I2) of pair b after a run: device pointer,
   pitch in floats; level 0 aliases the caller's input */
ICA_API int ica_plan_get_level_device(ica_plan* plan, int32_t which, int32_t pair, int32_t scale,
                              const float** ptr_out, int32_t* pitch_out);
/* number of kernels the last run launched (bench.py's gpu_launches) */
ICA_API int64_t ica_plan_last_launch_count(const ica_plan* plan);
/* Time (ms) spent in the per-iteration kernel / pyramid kernels during the last run.  enable = 1: CUDA events
   around the pyramid launches, the iterate kernel is timed by device-side %globaltimer stamps (first block start
   to last block end of every launch) accumulated by the solve kernel, so it also works inside the graph loop;
   enable = 2: host-driven loop with a CUDA-event pair around every iterate launch (cross-check). */
ICA_API int ica_plan_enable_timing(ica_plan* plan, int32_t enable);
ICA_API int ica_plan_get_timing(ica_plan* plan, float* iterate_ms, int32_t* iterate_launches,
                        float* pyramid_ms, int32_t* pyramid_launches);

/* ---- stateless entry points (helper API of the reference; also the parity-test hooks) ------ */
/* bicubic_interpolation_skimage (src/bicubic_interpolation.py:154-206): order-3 warp of
   `image` [H][W][C] by the 3x3 row-major `matrix` (output (col,row) -> input), NaN where the
   4x4 footprint leaves the image, clipped to the image's [min,max].  Host buffers. */
ICA_API int ica_warp_host(const float* image, int32_t height, int32_t width, int32_t channels,
                  const double* matrix9, float* out);
/* skimage.transform.rescale as called at ica.py:333-336: one pyramid level.  out has shape
   zoom_size(height, width, nu) x C; out_h/out_w receive it. */
ICA_API int ica_rescale_host(const float* image, int32_t height, int32_t width, int32_t channels,
                     double nu, float* out, int32_t* out_h, int32_t* out_w);
/* The banded 1-D operator the pyramid kernels apply along one axis (host computation, no GPU needed):
   out[o] = sum_k weights[o*taps + k] * in[start[o] + k] reproduces, per axis, skimage.transform.rescale's
   Gaussian -> 12-sample zero pad -> cubic-spline prefilter -> spline evaluation (ica.py:333-336).
   weights_out needs n_out * 64 floats at most; fast_range_out[3] = {lo, hi, s0} of the uniform rows. */
ICA_API int ica_resample_operator(int32_t n_in, int32_t n_out, int32_t* taps_out, int32_t* start_out,
                                  float* weights_out, int32_t weights_capacity, int32_t* fast_range_out);
/* zoom.zoom_out (src/zoom.py:29-60, the IPOL-style level: Gaussian sigma = 0.6 sqrt(1/f^2 - 1) with scipy's reflect
   extension, then cubic-spline map_coordinates(mode='nearest') at o / f).  The reference's function is dead code that
   raises on current scipy, so parity is UNPINNED (oracle restates the algorithm it spells out).
   ica_zoom_out_operator: the 1-D banded operator (host only, no device); ica_zoom_out_host: one image, float32
   [H][W][C] -> [round(H f)][round(W f)][C]. */
ICA_API int ica_zoom_out_operator(int32_t n_in, double factor, int32_t* n_out, int32_t* taps_out, int32_t* start_out,
                                  float* weights_out, int32_t weights_capacity);
ICA_API int ica_zoom_out_host(const float* image, int32_t height, int32_t width, int32_t channels, double factor,
                              float* out, int32_t* out_h, int32_t* out_w);
/* Synthetic pairs with a known ground truth, generated on the device (the reference fabricates its test
   pairs with transformation.transform_image, src/transformation.py:266-318, inside the notebooks): smooth random
   texture, I2 = its centre crop, I1(x) = texture(x'(x; p_gt)) + noise (+ an occluding square of uniform noise), values
   in [0, 255], optionally rounded to 8-bit.  Counter-based randomness: a value depends on (seed, pair_offset + pair,
   position) only.  I1/I2: float32 [B][H][W][C] device buffers; ttypes [B], p_gt [B][8], occ_xy [B][2] (or NULL) host. */
ICA_API int ica_generate_pairs_device(float* I1_dev, float* I2_dev, int32_t batch, int32_t height, int32_t width,
                                      int32_t channels, const int32_t* ttypes, const double* p_gt, const int32_t* occ_xy,
                                      int32_t occ_side, uint64_t seed, int32_t pair_offset, int32_t margin,
                                      double noise_sigma, int32_t quantize, void* stream);
/* zoom.zoom_size (src/zoom.py:8-22), round-half-to-even */
ICA_API int ica_zoom_size(int32_t nx, int32_t ny, double factor, int32_t* nxx, int32_t* nyy);
/* Gradient of I1 + frame (ica.py:81-93): Ix, Iy float32 [H][W][C]; NaN on the frame */
ICA_API int ica_gradient_host(const float* image, int32_t height, int32_t width, int32_t channels,
                      int32_t delta, int32_t nanifoutside, float* Ix, float* Iy);
/* One evaluation of the fused per-iteration kernel on one pair at one scale: returns the
   robust-weighted Hessian H[n*n] and vector b[n] (de.hessian[_robust] + io.independent_vector[_robust],
   src/derivatives.py:73-107, src/image_optimisation.py:82-143) for parameters p. */
ICA_API int ica_hessian_b_host(const float* I1, const float* I2, int32_t height, int32_t width,
                       int32_t channels, int32_t gray_as_rgb, int32_t transform_type,
                       const double* p, int32_t robust_type, double lambda_, int32_t delta,
                       int32_t nanifoutside, double* H_out, double* b_out);
/* host-side scalar algebra, same source as the device epilogue (csrc/ica_transform.cuh) */
ICA_API int ica_nparams(int32_t transform_type);                                  /* tr.py:15-32 */
ICA_API int ica_params2matrix(const double* p, int32_t transform_type, double* m9); /* tr.py:188-236 */
ICA_API int ica_update_transform(double* p_inout, const double* dp, int32_t transform_type); /* tr.py:36-141 */
ICA_API int ica_zoom_in_parameters(const double* p, int32_t transform_type, double nx, double ny,
                           double nxx, double nyy, double* p_out);         /* zoom.py:62-125 */
/* de.inverse_hessian (src/derivatives.py:110-130): LU inverse with partial pivoting, zero
   matrix when exactly singular; n <= 8 */
ICA_API int ica_inverse_hessian(const double* H, int32_t n, double* H_inv);

/* ---- helper API on materialised float64 arrays (what callers of the reference's helper modules import).
   The drivers never materialise these arrays; the entry points exist for callers of the helpers.  Host buffers. */
/* io.rhop (src/image_optimisation.py:17-53), element-wise over `count` values of t2 */
ICA_API int ica_rhop_host(const double* t2, int64_t count, double lambda_, int32_t robust_type, double* out);
/* io.robust_error_function (io.py:56-79): DI [H][W][C] -> rho [H][W] */
ICA_API int ica_robust_error_host(const double* DI, int32_t height, int32_t width, int32_t channels, double lambda_,
                                  int32_t robust_type, double* rho_out);
/* io.steepest_descent_images (io.py:158-194): Ix, Iy [H][W][C], J [H][W][2n] -> DIJ [H][W][C][n] */
ICA_API int ica_steepest_descent_host(const double* Ix, const double* Iy, const double* J, int32_t height, int32_t width,
                                      int32_t channels, int32_t nparams, double* DIJ_out);
/* de.hessian / de.hessian_robust (derivatives.py:73-107) when DI == NULL: out = H [n][n];
   io.independent_vector / _robust (io.py:82-143) when DI is given: out = b [n].  rho == NULL: unweighted. */
ICA_API int ica_dij_reduce_host(const double* DIJ, const double* DI, const double* rho, int32_t height, int32_t width,
                                int32_t channels, int32_t nparams, double* out);
/* tr.transform_image (transformation.py:266-318): skimage warp, order 1, cval 0, clip; matrix9 maps output (col,row) to
   input (the caller passes the inverse of the transform it wants to apply, as the reference does with tform.inverse) */
ICA_API int ica_transform_image_host(const double* image, int32_t height, int32_t width, int32_t channels,
                                     const double* matrix9, double* out);

/* out = A_y * image * A_x^T for caller-supplied banded operators: output row o of axis y reads input rows
   ystart[o] .. ystart[o]+ytaps-1 with weights yweights[o*ytaps ..]; same along x (at most 128 taps).  Optionally clipped to
   the input's [min,max].  Runs on the pyramid kernels; used by the mirror of zm.zoom_out (zoom.py:29-60). */
ICA_API int ica_apply_operators_host(const float* image, int32_t height, int32_t width, int32_t channels,
                                     const int32_t* ystart, const float* yweights, int32_t ytaps, int32_t ny_out,
                                     const int32_t* xstart, const float* xweights, int32_t xtaps, int32_t nx_out,
                                     int32_t clip_to_input_range, float* out);
/* bi.bicubic_interpolation_image (bicubic_interpolation.py:121-152): the IPOL-style warp -- the model is selected by
   the number of parameters (tr.project, transformation.py:144-186), NaN (nanifoutside) or 0 within `delta` of the border
   of the projected domain, Catmull-Rom with clamped (Neumann) neighbours, no clipping.  Not used by the drivers. */
ICA_API int ica_warp_ipol_host(const double* image, int32_t height, int32_t width, int32_t channels, const double* params,
                               int32_t nparams, int32_t nanifoutside, int32_t delta, double* out);

#ifdef __cplusplus
}
#endif
#endif /* ICA_B200_H */
