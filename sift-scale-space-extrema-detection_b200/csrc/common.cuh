// common.cuh -- shared declarations between the engine (engine.cu) and the kernel
// translation units.  sm_100a only; no fallbacks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sift_b200.h"

#define SIFT_MAX_OCTAVES 12
#define SIFT_MAX_LEVELS 12   // scalesPerOctave + 3 <= 12

// Device-side description of one octave of the pyramid held by a context.
// All planes are fp32 with a common pitch (in elements, multiple of 32 so that
// every row starts on a 128-byte line); the seed (level 0 of octaves >= 1,
// background.js:114-130) is additionally kept unrounded in fp64 so that the next
// octave is blurred from the same numbers the reference blurs from.
struct OctaveDev {
  int w, h;            // columns, rows of this octave
  int pitch;           // elements per row of the fp32 planes
  int nlev;            // Gaussian levels (spo + 3)
  float *gauss[SIFT_MAX_LEVELS];
  float *dog[SIFT_MAX_LEVELS];
  double *seed64;      // h * w dense, octaves >= 1 (nullptr for octave 0)
  // ---- mosaic strips: this octave image is rows [y_top, y_top + h) of a taller global
  // octave of gh rows; rows [own0, own1) (local) are owned by this strip, the rest is halo.  A whole image
  // is the strip y_top = 0, gh = h, own = [0, h), seed_off = 0.
  int y_top, gh, own0, own1;
  int seed_off;        // local row (y >> 1) + seed_off of the next octave receives this octave's even row y
  // local rows [valid0, valid1) hold the whole image's values: a level row within the largest radius of a strip
  // edge that is not an image edge was blurred against the strip's clamp, not against the neighbour's rows
  int valid0, valid1;
};

// Per-level blur description (host computed, background.js:156-177 + sift.js:38).
struct LevelPlan {
  double blurLevel;    // target sigma (scale_space[o][s].blurLevel)
  double offsetSigma;  // sigma of the kernel applied to the octave base (0 for a seed level)
  int radius;          // round(3 * offsetSigma), sift.js:38
  int woff;            // offset of this level's padded 1D weights in the weight buffer
};

// Counter block in device memory (one per context).
struct Counters {
  int n_cand;          // candidates appended
  int n_low;           // low-contrast extrema (when counted)
  int n_kp;            // keypoints appended
  int outcomes[8];     // refine outcomes, index = REFINE_* below
  int n_left_strip;    // refinement walks that left the rows this strip holds (not the image): handed out as sift_walk
  int pad[4];
};

enum {
  REFINE_ACCEPTED = 0, REFINE_LOW_CONTRAST = 1, REFINE_EDGE = 2, REFINE_LEFT_SCALE = 3,
  REFINE_LEFT_ROWS = 4, REFINE_LEFT_COLS = 5, REFINE_NO_CONVERGENCE = 6, REFINE_SINGULAR = 7
};

// Number of zero weights appended after each level's 2R+1 taps so that the
// register-rotating inner loops can run past the end branch-free.
#define SIFT_WPAD 16

// ---- launchers implemented in the kernel translation units --------------------
// blur_generic.cu
void launch_hblur(cudaStream_t st, const void *src, int dtype, size_t src_pitch_bytes, int src_w, int src_h,
                  int upsample, int w, int hrows, const double *d_weights, const LevelPlan *plans, int first_level,
                  int nlev, double *const *T /* nlev device planes hrows*w */, double **d_Tptrs);
void launch_vblur(cudaStream_t st, int upsample, const OctaveDev &oct, const double *d_weights,
                  const LevelPlan *plans, int first_level, double *const *d_Tptrs_dev,
                  const OctaveDev *next /* may be null */, int spo, int keep_gauss);
void launch_blur_plane_f64(cudaStream_t st, const double *d_in, int rows, int cols, double *d_tmp, double *d_out,
                           const double *d_w, int radius, int x1, int y1, int x2, int y2);
void launch_subtract_f64(cudaStream_t st, const double *a, const double *b, double *out, int cols,
                         int x1, int y1, int x2, int y2);
void launch_resize_f64(cudaStream_t st, const double *in, int rows, int cols, double rate, double *out,
                       int orows, int ocols);
void launch_seed_to_f32(cudaStream_t st, const double *seed, int w, int h, float *dst, int pitch);

// blur_sep.cu: pass A (along x) -> fp64 T^T planes -> pass B (along y), all levels of an octave per launch
bool sep_supported(const LevelPlan *plans, int first_level, int nlev, int w, int h);
size_t sep_t_elems(int w, int h, int trows, const LevelPlan *plans, int first_level, int nlev);
size_t sep_tma_map_bytes(int n_levels);
int sep_tma_build_maps(const LevelPlan *plans, int first_level, int nlev, double *tbase, int w, int h, int trows, void *h_maps);
void launch_sep_pass_a(cudaStream_t st, const void *src, int dtype, size_t src_pitch_bytes, int src_w, int upsample,
                       int w, int hrows, int h, const double *d_weights, const LevelPlan *plans, int first_level,
                       int nlev, double *tbase);
bool sep_pass_b_uses_tma(const void *d_tmaps, int upsample);
void launch_sep_pass_b(cudaStream_t st, int upsample, const OctaveDev &oct, const double *d_weights,
                       const LevelPlan *plans, int first_level, double *tbase, int hrows, const OctaveDev *next,
                       int spo, int keep_gauss, const void *d_tmaps);

// blur_fused.cu
void fused0_merge_taps(const double *w, int R, double *out /* fused0_taps_per_level() doubles */);
int fused0_taps_per_level(void);
bool fused0_supported(const LevelPlan *plans, int nlev);
void launch_fused_octave0(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                          const OctaveDev &oct, const OctaveDev *next, const double *d_weights,
                          const LevelPlan *plans, int poly_woff, int nlev, int spo, int keep_gauss,
                          const double *d_u8lut);

// blur_mma.cu: the blur as banded-Toeplitz products on the fp64 matrix instruction (DMMA.8x8x4)
bool mma0_supported(const LevelPlan *plans, int nlev);
int mma0_frag_doubles(int nlev);
void mma0_build_frags(const double *merged, int R, double *out /* mma0_frag_doubles(1) */);
bool launch_oct0_mma(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                     const OctaveDev &oct, const OctaveDev *next, const double *d_frags, const LevelPlan *plans,
                     int nlev, int spo, int keep_gauss, int tile_row0 = 0, int tile_rows = -1);
int mma0_tile_rows(const OctaveDev &oct, int src_h);   // tile rows (32 source rows each) of a whole launch

// octaves >= 1 as two DMMA passes (row-major fp64 intermediate)
bool mma_sep_supported(const LevelPlan *plans, int nlev, int w, int h);
size_t mma_sep_t_elems(const LevelPlan *plans, int nlev, int w, int h);
void launch_mma_sep(cudaStream_t st, const OctaveDev &oct, const OctaveDev *next, const double *d_weights,
                    const LevelPlan *plans, double *tbase, int spo, int keep_gauss);

// blur_oct0.cu: octave 0, second generation (column pass first, row pass last; TMA source tile)
bool oct0_v2_supported(const LevelPlan *plans, int nlev);
size_t oct0_out_map_bytes(int nlev);
bool oct0_build_out_maps(const OctaveDev &oct, int nlev, void *h_maps /* oct0_out_map_bytes(nlev) */);
void launch_oct0_v2(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                    const OctaveDev &oct, const OctaveDev *next, const double *d_weights, const LevelPlan *plans,
                    int poly_woff, int nlev, int spo, int keep_gauss, const void *d_out_maps);

// blur_oct0s.cu: octave 0, small-CTA form (64 x 32 output tiles, <= 64 registers, four CTAs per SM)
bool oct0_small_supported(const LevelPlan *plans, int nlev);
void launch_oct0_small(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                       const OctaveDev &oct, const OctaveDev *next, const double *d_weights, const LevelPlan *plans,
                       int poly_woff, int nlev, int spo, int keep_gauss);

// blur_oct0p.cu: octave 0 as two polyphase passes over row bands, intermediate in L2
bool oct0p_supported(const LevelPlan *plans, int nlev);
int oct0p_band_rows(int src_w, int src_h, int nlev);
size_t oct0p_t_bytes(int src_w, int src_h, int nlev);
size_t oct0p_map_bytes(int nlev);
bool oct0p_build_maps(double *tbase, int src_w, int src_h, int nlev, void *h_maps);
int launch_oct0p(cudaStream_t st, const void *src, int dtype, size_t src_pitch, int src_w, int src_h,
                 const OctaveDev &oct, const OctaveDev *next, const double *d_weights, const LevelPlan *plans,
                 int poly_woff, int nlev, int spo, int keep_gauss, double *tbase, const void *d_maps);

// scan.cu
void launch_scan_all(cudaStream_t st, const OctaveDev *h_octs, const OctaveDev *d_octs, int n_oct, int spo,
                     double pix_threshold, int count_low, sift_candidate *cand, int cand_cap, sift_candidate *low,
                     int low_cap, Counters *ctr);
// TMA-tiled scan (one tensor map per octave and DoG level, built per lane when its pyramid is laid out)
size_t scan_tma_map_bytes(int n_oct, int ndog);
bool scan_tma_supported(int ndog);
int scan_tma_build_maps(const OctaveDev *octs, int n_oct, int ndog, void *h_maps);
void launch_scan_tma(cudaStream_t st, const OctaveDev *h_octs, const void *d_maps, int n_oct, int spo,
                     double pix_threshold, int count_low, sift_candidate *cand, int cand_cap, sift_candidate *low,
                     int low_cap, Counters *ctr);
void launch_scan_f64(cudaStream_t st, const double *d0, const double *d1, const double *d2, int rows, int cols,
                     double pix_threshold, int32_t *cand_xy, double *cand_val, int cand_cap,
                     int32_t *low_xy, double *low_val, int low_cap, int *counts /* [2] */);

// preview.cu: display products (RGBA8) of one level
int launch_preview(cudaStream_t st, const float *src, int w, int h, size_t pitch, int mode, double coefficient,
                   void *scratch32, void *d_out);

// order.cu: device-side ordering of keypoint records (reference order: octave, scale, row, column)
size_t order_scratch_bytes(int n_sort);
int launch_order_keypoints(cudaStream_t st, const sift_keypoint *kp, const Counters *ctr, int n_sort, void *scratch,
                           size_t scratch_bytes, sift_keypoint *out, int cap, int *d_count);

// refine.cu (compiled with --fmad=false: the reference never fuses multiply-add)
struct RefineParams {
  int spo, ndog, max_iter;
  double offset_bound, contrast_thr, edge_thr, min_blur, min_interpixel;
};
// walks / walk_cap: where walks that leave a strip's rows are recorded (slot = Counters::n_left_strip); may be NULL
void launch_refine(cudaStream_t st, const OctaveDev *d_octs, int n_octs, const sift_candidate *cand,
                   const int *d_ncand, int n_cand_host /* -1: read d_ncand */, int cand_cap, RefineParams rp,
                   sift_keypoint *out, int cap, Counters *ctr, sift_walk *walks, int walk_cap);
void launch_refine_resume(cudaStream_t st, const OctaveDev *d_octs, const sift_walk *in, int n, RefineParams rp,
                          sift_keypoint *out, int cap, Counters *ctr, sift_walk *walks, int walk_cap);
void launch_grad_hess_f64(cudaStream_t st, const double *dm, const double *dc, const double *dp, int cols,
                          int m, int n, double *out12);
