/*
 * sift_b200.h -- C ABI of the B200-native SIFT detection engine (libsift_b200.so).
 *
 * This is the drop-in boundary for the detection path of
 * bingjetli/sift-scale-space-extrema-detection: what a Node N-API addon (see
 * INTEGRATION.md), the Python host mirror (ctypes) and the parity tests bind.
 * Plain C types only; no C++ types, exceptions or torch types cross it.
 *
 * Reference interfaces replaced (file:line into the reference tree):
 *   coarse stage functions  background.js:71   computeGaussianScaleSpace
 *                           background.js:258  computeDifferenceOfGaussians
 *                           background.js:359  findCandidateKeypoints
 *                           background.js:455  refineCandidateKeypoints
 *   fine step functions     src/sift.js:72     SIFT_blurMatrix2DChunk
 *                           src/sift.js:154    SIFT_subtractMatrix2DChunk
 *                           src/sift.js:212    SIFT_findExtremas
 *                           src/sift.js:333    SIFT_generateGradientVector
 *                           src/sift.js:377    SIFT_generateHessianMatrix
 *   helpers on the path     src/matrix2d.js:112 Matrix2D_linearResize
 *                           src/matrix2d.js:464 Matrix2D_get3x3Inverse
 *   transport replaced      src/worker.js:29-98 (postMessage requests) and
 *                           background.js:14-50 (onmessage switch)
 *
 * Conventions
 *   - every function returns an int status (SIFT_OK == 0); sift_last_error()
 *     returns the text of the last failure on that context.
 *   - the caller owns every host buffer; the library never retains a caller
 *     pointer past the call.  Device pyramids are owned by the context and
 *     stay valid until the next build/detect call on it.
 *   - a context is single-threaded / non-reentrant: one context per GPU per
 *     host thread (the reference's worker also handles one message at a time,
 *     background.js:14-50).
 *   - there is NO CPU fallback: without an sm_100 device sift_create fails.
 *   - images are row-major; Matrix2D m[i][j] (row i, column j) == p[i*cols + j].
 */
#ifndef SIFT_B200_H
#define SIFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SIFT_API __declspec(dllexport)
#else
#define SIFT_API __attribute__((visibility("default")))
#endif

/* ----------------------------------------------------------------- status */
enum {
  SIFT_OK = 0,
  SIFT_ERR_BAD_ARGS = 1,       /* null pointer, non-positive size, sigma schedule not realisable */
  SIFT_ERR_CUDA = 2,           /* a CUDA runtime call failed (text in sift_last_error) */
  SIFT_ERR_CAPACITY = 3,       /* output capacity too small; *n_out holds the required count */
  SIFT_ERR_UNSUPPORTED = 4,    /* size / parameter outside what the engine supports */
  SIFT_ERR_NO_DEVICE = 5,      /* no CUDA device of compute capability 10.x */
  SIFT_ERR_STATE = 6           /* stage called out of order (no pyramid built yet) */
};

/* ------------------------------------------------------------ pixel types */
enum {
  SIFT_U8 = 0,   /* grey bytes; float image = v / 255.0 (image-utils.js:114) */
  SIFT_F32 = 1,  /* float image already in [0,1] */
  SIFT_F64 = 2,  /* double image already in [0,1] (the reference's Matrix2D numbers) */
  SIFT_RGBA8 = 3 /* ImageData bytes; grey = (0.299R + 0.587G + 0.114B) / 255.0 (image-utils.js:107-114) */
};

enum { SIFT_LEVEL_GAUSSIAN = 0, SIFT_LEVEL_DOG = 1 };

/* ------------------------------------------------------------- parameters */
/* Names follow the reference's request fields (worker.js:40-48, 90-98); the
 * constants the reference hard-codes are defaulted fields. */
typedef struct sift_params {
  int32_t numberOfOctaves;       /* worker.js:33  default 5 */
  int32_t scalesPerOctave;       /* worker.js:34  default 3 */
  double minBlurLevel;           /* worker.js:35  default 0.8 */
  double assumedBlur;            /* worker.js:36  default 0.5 */
  double contrastThreshold;      /* sift.js:285 / background.js:572  0.015 */
  double preFilterFactor;        /* sift.js:293  0.8 */
  double edgeRatio;              /* background.js:598  10 */
  int32_t maxIterations;         /* background.js:480  5 */
  int32_t reserved0;
  double offsetBound;            /* background.js:558  0.6 */
  double minInterpixelDistance;  /* background.js:461  0.5 */
} sift_params;

/* ---------------------------------------------------------------- records */
/* sift.js:274-278 extremum {x, y, value} plus where it was found
 * (background.js:433-436 groups them per octave / scaleLevel). */
typedef struct sift_candidate {
  int32_t octave;
  int32_t scaleLevel;
  int32_t x;      /* column */
  int32_t y;      /* row */
  float value;    /* DoG value at the extremum */
  int32_t reserved0;
} sift_candidate; /* 24 bytes */

/* background.js:619-628 refined keypoint record (octave, scaleLevel, localX,
 * localY, absoluteSigma, absoluteX, absoluteY, interpolatedValue) plus the
 * interpolation offsets in [s, m(row), n(col)] order, the DoG value of the
 * originating extremum, and that extremum's position (the reference's output
 * order is the candidate order: octave, scale, y, x). */
typedef struct sift_keypoint {
  int32_t octave;
  int32_t scaleLevel;   /* post-move s */
  int32_t localX;       /* post-move n (column) */
  int32_t localY;       /* post-move m (row) */
  double absoluteSigma;
  double absoluteX;
  double absoluteY;
  double interpolatedValue;
  float offset[3];      /* alpha: scale, row, column */
  float dogValue;       /* extrema.value (background.js:565) */
  int32_t candScale;
  int32_t candX;
  int32_t candY;
  int32_t iterations;   /* quadratic-fit iterations before acceptance (0-based) */
} sift_keypoint; /* 80 bytes */

/* Per-call counters (the reference only console.logs these, background.js:581-672). */
typedef struct sift_stats {
  int32_t candidates;       /* extrema with abs(value) >= preFilterFactor * threshold */
  int32_t lowContrastExtrema; /* extrema below it; -1 when not counted */
  int32_t keypoints;        /* accepted */
  int32_t rejLowContrast;   /* background.js:577 */
  int32_t rejEdge;          /* background.js:599 */
  int32_t rejLeftScale;     /* background.js:644 */
  int32_t rejLeftRows;      /* background.js:651 */
  int32_t rejLeftCols;      /* background.js:658 */
  int32_t rejNoConvergence; /* loop ran out, background.js:480 */
  int32_t rejSingular;      /* abs(det) < Number.EPSILON (matrix2d.js:482): the reference throws, we discard */
  float msDevice;           /* CUDA-event time of the device work of this call */
  int32_t kernelLaunches;   /* kernels launched by this call */
  int32_t leftStrip;        /* mosaic strips only: walks that jumped to a row this strip does not hold (not out of the
                             * image).  They are NOT counted in any rej* field here: each is handed out as a sift_walk
                             * (sift_strip_escaped) and its outcome is counted by the strip that ends it */
} sift_stats;

typedef struct sift_ctx sift_ctx;

/* ------------------------------------------------------- context / errors */
SIFT_API int sift_create(int device, sift_ctx **out);
SIFT_API void sift_destroy(sift_ctx *ctx);
SIFT_API const char *sift_last_error(const sift_ctx *ctx);   /* ctx may be NULL: last create error */
SIFT_API const char *sift_version(void);
SIFT_API void sift_default_params(sift_params *p);           /* worker.js:33-37 defaults */
SIFT_API int sift_synchronize(sift_ctx *ctx);
/* The public CUDA stream (cudaStream_t) of the context.  Device-resident calls fork from it (they see
 * everything queued on it before the call) and run on internal lanes -- one stream + pyramid per frame in
 * flight -- so consecutive frames overlap.  sift_flush() makes the public stream wait for all lanes
 * (without blocking the host): queue it before consuming device results or recording an event there. */
SIFT_API void *sift_stream(sift_ctx *ctx);
SIFT_API int sift_flush(sift_ctx *ctx);
/* Keypoint-only callers do not read the Gaussian levels back: keep == 0 stops the blur kernels from writing
 * them (the DoG levels and the seeds are unaffected; sift_get_level(GAUSSIAN) then returns stale data for
 * levels >= 1).  Default 1: every level of the pyramid is materialised once, like the reference's reply. */
SIFT_API int sift_set_keep_gaussian(sift_ctx *ctx, int keep);
/* Frames in flight for sift_detect_device / sift_detect_batch: 1..8, or 0 (default; env SIFT_B200_LANES) = chosen
 * from the frame size -- 8 for frames below 6 Mpixel, 6 above (fewer when the lanes' pyramids would exceed ~32 GB). */
SIFT_API int sift_set_lanes(sift_ctx *ctx, int n_lanes);
SIFT_API int64_t sift_kernel_launches(const sift_ctx *ctx); /* running total for the context */
/* Generation of the pyramid the stage calls read (sift_get_level, sift_find_candidates, sift_refine ...): changes
 * whenever any call rebuilds, replaces or invalidates it (sift_build_scale_space, sift_detect*, sift_set_level,
 * sift_strip_begin, a plan change).  A host that caches "the context still holds the pyramid of reply X" -- as the
 * reference's main thread holds gaussian_scale_space / difference_of_gaussians (main.js:31-32) -- compares this. */
SIFT_API uint64_t sift_pyramid_serial(const sift_ctx *ctx);
/* Test hook: fill every device buffer of the context (planes, seeds, intermediates, record buffers) with `byte` and
 * invalidate the pyramid.  Results of later calls must not depend on the pattern (tests/test_memory_hygiene.py). */
SIFT_API int sift_debug_poison(sift_ctx *ctx, int byte);
/* Running totals of the bytes the detect calls (sift_detect, sift_detect_batch, stage uploads) copied host -> device
 * (images) and device -> host (counters + keypoint records), counted where the copies are issued. */
SIFT_API void sift_transfer_bytes(const sift_ctx *ctx, uint64_t *h2d, uint64_t *d2h);

/* Per-kernel-class device timing (CUDA events around each launch group on sift_stream).
 * Off by default; bench.py turns it on for a separate instrumented pass. */
enum {
  SIFT_PROF_BLUR_OCT0 = 0,  /* octave 0: upsample + blur + DoG + seed */
  SIFT_PROF_BLUR_OCT1 = 1,  /* octave 1 */
  SIFT_PROF_BLUR_HIGH = 2,  /* octaves >= 2 */
  SIFT_PROF_SCAN = 3,       /* extremum scan + compaction */
  SIFT_PROF_REFINE = 4,     /* refinement */
  SIFT_PROF_NKINDS = 5
};
SIFT_API int sift_set_profiling(sift_ctx *ctx, int enabled);
/* Sums (ms) and counts of the spans recorded since the last call; synchronises the stream. */
SIFT_API int sift_get_profile(sift_ctx *ctx, float *ms_by_kind, int *launch_groups_by_kind, int n_kinds);

/* ----------------------------------------------------- fused detect (hot) */
/* Whole path: 2x upsample, pyramid, DoG, extrema, refinement; HOST buffers in
 * and out (copies inside).  Keypoints come back in the reference's order.
 * pitch_bytes == 0 means dense rows.  out may be NULL with cap 0 to count. */
SIFT_API int sift_detect(sift_ctx *ctx, const void *image, int dtype, int width, int height,
                         size_t pitch_bytes, const sift_params *params,
                         sift_keypoint *out, int cap, int *n_out, sift_stats *stats);

/* Same, image already in device memory; keypoints stay in device memory
 * (d_out; in the reference's order when `ordered` != 0, which adds a device-side radix sort, else in
 * completion order) and *d_count (device int) receives the count.  Overflow is reported on the device, never
 * silently truncated: when the context's candidate list or `cap` (ordered: min(cap, the lane's record buffer))
 * was too small, *d_count = -(capacity that would have sufficed) and the records written are an arbitrary
 * subset that must not be used -- call again with a larger cap / through sift_detect, which grows the buffers.
 * Asynchronous: ordered after what is already queued on sift_stream(ctx);
 * sift_flush() / sift_synchronize() order the results before later work on that stream. */
SIFT_API int sift_detect_device(sift_ctx *ctx, const void *d_image, int dtype, int width, int height,
                                size_t pitch_bytes, const sift_params *params,
                                sift_keypoint *d_out, int cap, int *d_count, int ordered);

/* n_images equally sized host images, image i at image + i*image_stride_bytes.
 * Keypoints of image i are out[offsets[i] .. offsets[i+1]) (offsets has
 * n_images+1 entries). */
SIFT_API int sift_detect_batch(sift_ctx *ctx, const void *images, int dtype, int width, int height,
                               size_t pitch_bytes, size_t image_stride_bytes, int n_images,
                               const sift_params *params, sift_keypoint *out, int cap,
                               int *offsets, sift_stats *stats);

/* -------------------------------------------------- coarse stage functions */
/* background.js:71  -- builds Gaussian levels (and, fused, the DoG levels) on the device. */
SIFT_API int sift_build_scale_space(sift_ctx *ctx, const void *image, int dtype, int width, int height,
                                    size_t pitch_bytes, const sift_params *params);
/* background.js:258 -- DoG of the pyramid held by the context (already formed
 * from the unrounded accumulators by sift_build_scale_space; this validates state). */
SIFT_API int sift_build_dog(sift_ctx *ctx);
/* background.js:359 -- candidates in reference order. low / n_low may be NULL.  params may be
 * NULL (the context's); only contrastThreshold and preFilterFactor are read from it. */
SIFT_API int sift_find_candidates(sift_ctx *ctx, const sift_params *params, sift_candidate *out, int cap,
                                  int *n_out, sift_candidate *low, int low_cap, int *n_low);
/* background.js:455 -- refine a caller-supplied candidate list against the DoG held by the context.
 * params may be NULL; only the threshold / position fields are read (contrastThreshold, edgeRatio,
 * maxIterations, offsetBound, minBlurLevel, minInterpixelDistance): the reference passes minBlurLevel
 * and minInterpixelDistance with this request (worker.js:90-98). */
SIFT_API int sift_refine(sift_ctx *ctx, const sift_params *params, const sift_candidate *cands, int n_cands,
                         sift_keypoint *out, int cap, int *n_out, sift_stats *stats);

/* Pyramid inspection (scale_space[o][s] = {blurLevel, image}, background.js:57-70). */
SIFT_API int sift_get_pyramid_info(const sift_ctx *ctx, int *octaves, int *levels_per_octave);
SIFT_API int sift_get_octave_size(const sift_ctx *ctx, int octave, int *width, int *height);
SIFT_API int sift_get_blur_level(const sift_ctx *ctx, int kind, int octave, int level, double *blur_level);
SIFT_API int sift_get_level(sift_ctx *ctx, int kind, int octave, int level, float *dst /* h*w dense */);
/* Display products of one level of the held pyramid: an RGBA8 image in ImageData layout
 * (w*h*4 bytes, R=G=B=grey, A=255), as the reference posts beside each stage result.
 *   SIFT_PREVIEW_GRAY     Math.round(v*255)                    ImageUtils_convertMatrix2DToImageData, image-utils.js:171-220
 *   SIFT_PREVIEW_SIGMOID  1/(1+exp(coefficient*(-v))) first    Matrix2D_sigmoidNormalize, matrix2d.js:148-156 (5: background.js:307)
 *   SIFT_PREVIEW_MINMAX   (v-min)/(max-min) of the level first Matrix2D_sampledNormalize, matrix2d.js:169-192 (background.js:336)
 * min_max (optional, 2 doubles) receives the level's min / max in MINMAX mode, {0, 1} otherwise. */
enum { SIFT_PREVIEW_GRAY = 0, SIFT_PREVIEW_SIGMOID = 1, SIFT_PREVIEW_MINMAX = 2 };
SIFT_API int sift_get_level_preview(sift_ctx *ctx, int kind, int octave, int level, int mode, double coefficient,
                                    unsigned char *rgba_out /* h*w*4 */, double *min_max);
/* Replace a DoG level of the held pyramid with caller data (used by the host mirror
 * when refineCandidateKeypoints / findCandidateKeypoints receive foreign matrices). */
SIFT_API int sift_set_pyramid_shape(sift_ctx *ctx, int width0, int height0, const sift_params *params);
SIFT_API int sift_set_level(sift_ctx *ctx, int kind, int octave, int level, const float *src);

/* ---------------------------------------------------------- mosaic strips */
/* A mosaic too large for one GPU is cut into horizontal strips of the octave-0 grid, one per GPU
 * (SURVEY.md 8e).  Every level of an octave is blurred from that octave's base image only
 * (background.js:185-190), so the only cross-strip dependency is the base (seed) image's halo: per octave a
 * strip needs `halo` = max kernel radius + margin rows of its neighbours' seed.  The engine computes one
 * octave at a time; between octaves the caller exchanges seed rows (NCCL / P2P) directly in device memory.
 * All row numbers below are rows of the GLOBAL octave grids. */
typedef struct sift_strip_layout {
  int32_t octaves;
  int32_t width[12], height[12];   /* global octave sizes */
  int32_t own0[12], own1[12];      /* rows this strip owns (its keypoints come from these) */
  int32_t top[12], bottom[12];     /* rows this strip holds: owned rows + halos, clipped to the image */
  int32_t halo[12];                /* rows needed beyond the owned range: max radius + margin, even */
} sift_strip_layout;
/* Host-only arithmetic (no device needed).  row0/row1: owned octave-0 rows, multiples of 2^(octaves-1)
 * (row1 may also be the image bottom, 2 * full_height).  margin: rows kept beyond the blur halo for the
 * scan and for refinement walks. */
SIFT_API int sift_strip_layout_compute(const sift_params *params, int full_width, int full_height, int row0,
                                       int row1, int margin, sift_strip_layout *out);
/* Lay the strip's pyramid out on the device (lane 0) and upload the source rows
 * [top[0]/2, (bottom[0]+1)/2) of the full image (`rows` points at the first of them).  The upload is asynchronous
 * (chunks on a copy stream; octave 0 runs band by band behind them): `rows` must stay valid until
 * sift_strip_octave(ctx, 0) has returned, and should be page-locked for the copy to overlap. */
SIFT_API int sift_strip_begin(sift_ctx *ctx, const sift_params *params, const sift_strip_layout *layout,
                              const void *rows, int dtype, size_t pitch_bytes);
/* Device pointer to the strip's fp64 seed image of octave >= 1: dense, width[octave] doubles per row, first row
 * = global row top[octave].  After sift_strip_octave(octave - 1) the OWNED rows are final; the caller fills
 * the halo rows from the neighbouring strips, then calls sift_strip_octave(octave). */
SIFT_API int sift_strip_seed(sift_ctx *ctx, int octave, double **d_seed);
/* The exchange step itself, for hosts that drive all strips from ONE process (a Node.js / C host with one context
 * per GPU): fills the halo rows of every strip's seed image of `octave` from the strips that own them, device to
 * device (cudaMemcpyPeerAsync: NVLink between GPUs), stream-ordered between the senders' sift_strip_octave(octave-1)
 * and the receivers' sift_strip_octave(octave); the host is not blocked.  Call it once per octave >= 1 with every
 * strip of the mosaic (any order).  One-process-per-GPU hosts exchange the rows of sift_strip_seed() themselves
 * (mosaic.py: torch.distributed / NCCL send-recv). */
SIFT_API int sift_mosaic_exchange(sift_ctx *const *ctxs, int n_strips, int octave);
/* Blur + DoG of one octave (0, 1, 2 ... in order) and the owned rows of the next octave's seed.  Synchronous. */
SIFT_API int sift_strip_octave(sift_ctx *ctx, int octave);
/* Scan + refine over all octaves: keypoints of the owned rows, global coordinates, reference order. */
SIFT_API int sift_strip_finish(sift_ctx *ctx, sift_keypoint *out, int cap, int *n_out, sift_stats *stats);
/* A refinement walk (background.js:480-668) may jump to a row this strip does not hold: the reference moves
 * the sample by Math.round(alpha) with no bound (background.js:638-640).  Such a walk is not decided here:
 * it is handed out as a sift_walk -- the state of the loop at the jump -- and continued by the strip that
 * owns the new row.  stats.leftStrip counts them; the walk's final outcome is counted by whoever ends it. */
typedef struct sift_walk {
  int32_t octave, scaleLevel, x, y; /* current sample: DoG scale s, column n, row m of the GLOBAL octave grid */
  int32_t iteration;                /* iterations already spent: the loop resumes at this index (< maxIterations) */
  float value;                      /* DoG value of the ORIGINAL candidate (background.js:565 keeps using it) */
  int32_t candScale, candX, candY;  /* the original candidate (ordering key of the record it may become) */
  int32_t reserved0;
} sift_walk;                        /* 40 bytes */
/* Walks that left this strip during the last sift_strip_finish / sift_strip_resume. */
SIFT_API int sift_strip_escaped(sift_ctx *ctx, sift_walk *out, int cap, int *n_out);
/* Continue walks on this strip's pyramid (call it on the strip that OWNS row walks[i].y; a walk whose row this
 * strip does not hold comes straight back through sift_strip_escaped).  Keypoints in reference order; stats
 * carries the outcomes of the walks ended here (candidates = 0). */
SIFT_API int sift_strip_resume(sift_ctx *ctx, const sift_walk *walks, int n, sift_keypoint *out, int cap, int *n_out,
                               sift_stats *stats);

/* ----------------------------------------------------- fine step functions */
/* src/sift.js:72  float64 in / out like Matrix2D; half-open chunk; writes only the chunk of output. */
SIFT_API int sift_blur_chunk(sift_ctx *ctx, const double *input, int rows, int cols, double *output,
                             double sigma, int x1, int y1, int x2, int y2);
/* src/sift.js:154  output = a - b over the chunk. */
SIFT_API int sift_subtract_chunk(sift_ctx *ctx, const double *a, const double *b, int rows, int cols,
                                 double *output, int x1, int y1, int x2, int y2);
/* src/sift.js:212  strict 26-neighbour extrema of the middle image, raster order. */
SIFT_API int sift_find_extremas(sift_ctx *ctx, const double *d0, const double *d1, const double *d2,
                                int rows, int cols, int scales_per_octave, double contrast_threshold,
                                double prefilter_factor,
                                int32_t *cand_xy, double *cand_value, int cand_cap, int *n_cand,
                                int32_t *low_xy, double *low_value, int low_cap, int *n_low);
/* src/sift.js:333 and :377 on three consecutive DoG images (s-1, s, s+1): g[3] in
 * [s, m, n] order and the symmetric 3x3 Hessian, row-major. */
SIFT_API int sift_gradient_hessian(sift_ctx *ctx, const double *dm, const double *dc, const double *dp,
                                   int rows, int cols, int m, int n, double g[3], double h[9]);
/* src/matrix2d.js:112  nearest-neighbour resample; out must hold the resized image
 * (sift_resize_dims gives its shape). */
SIFT_API int sift_resize_dims(int rows, int cols, double sampling_rate, int *out_rows, int *out_cols);
SIFT_API int sift_linear_resize(sift_ctx *ctx, const double *input, int rows, int cols,
                                double sampling_rate, double *output);

#ifdef __cplusplus
}
#endif
#endif /* SIFT_B200_H */
