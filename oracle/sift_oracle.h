/* sift_oracle.h -- TEST INFRASTRUCTURE ONLY (see sift_oracle.c). */
#ifndef SIFT_ORACLE_H
#define SIFT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_OCTAVES 12
#define ORACLE_MAX_LEVELS 12

typedef struct { int x, y; double value; } oracle_extremum;               /* sift.js:274-278 */
typedef struct { int octave, scale, x, y; double value; } oracle_candidate; /* background.js:433-436 flattened */

/* background.js:619-628 plus the offsets / DoG value / originating candidate. */
typedef struct {
  int octave, scaleLevel, localX, localY;
  double absoluteSigma, absoluteX, absoluteY, interpolatedValue;
  double offset[3];          /* alpha in [s, m, n] order */
  double dogValue;           /* extrema.value of the originating candidate */
  int candScale, candX, candY, iterations;
} oracle_keypoint;

typedef struct {
  int numberOfOctaves, scalesPerOctave;      /* worker.js:33-34 */
  double minBlurLevel, assumedBlur;          /* worker.js:35-36 */
  double contrastThreshold;                  /* 0.015: sift.js:285, background.js:572 */
  double preFilterFactor;                    /* 0.8: sift.js:293 */
  double edgeRatio;                          /* 10: background.js:598 */
  int maxIterations;                         /* 5: background.js:480 */
  double offsetBound;                        /* 0.6: background.js:558 */
  double minInterpixelDistance;              /* 0.5: background.js:461 */
} oracle_params;

typedef struct {
  int octaves, levels, spo;
  int rows[ORACLE_MAX_OCTAVES], cols[ORACLE_MAX_OCTAVES];
  double *gauss[ORACLE_MAX_OCTAVES][ORACLE_MAX_LEVELS];
  double blur[ORACLE_MAX_OCTAVES][ORACLE_MAX_LEVELS];
  double offset_sigma[ORACLE_MAX_OCTAVES][ORACLE_MAX_LEVELS];
  double *dog[ORACLE_MAX_OCTAVES][ORACLE_MAX_LEVELS];
  double dog_blur[ORACLE_MAX_OCTAVES][ORACLE_MAX_LEVELS];
} oracle_pyramid;

enum {
  ORACLE_REFINE_ACCEPTED = 0,
  ORACLE_REFINE_LOW_CONTRAST = 1,
  ORACLE_REFINE_EDGE = 2,
  ORACLE_REFINE_LEFT_SCALE = 3,
  ORACLE_REFINE_LEFT_ROWS = 4,
  ORACLE_REFINE_LEFT_COLS = 5,
  ORACLE_REFINE_NO_CONVERGENCE = 6,
  ORACLE_REFINE_SINGULAR = 7,
  ORACLE_REFINE_NOUTCOMES = 8
};

double oracle_js_round(double x);
enum { ORACLE_PREVIEW_GRAY = 0, ORACLE_PREVIEW_SIGMOID = 1, ORACLE_PREVIEW_MINMAX = 2 };
int oracle_preview(const double *m, int rows, int cols, int mode, double coefficient,
                   unsigned char *rgba, double *min_max);
int oracle_linear_resize_dims(int rows, int cols, double rate, int *orows, int *ocols);
int oracle_linear_resize(const double *in, int rows, int cols, double rate, double *out);
int oracle_kernel_radius(double sigma);
int oracle_build_gaussian_kernel(double sigma, double *kernel);
int oracle_blur_chunk(const double *input, int rows, int cols, double *output,
                      double sigma, int x1, int y1, int x2, int y2);
int oracle_blur_image(const double *input, int rows, int cols, double *output, double sigma);
int oracle_blur_image_separable(const double *input, int rows, int cols, double *output, double sigma);
int oracle_subtract_chunk(const double *a, const double *b, int rows, int cols, double *output,
                          int x1, int y1, int x2, int y2);
double oracle_contrast_threshold(int spo, double contrast);
int oracle_find_extremas(const double *d0, const double *d1, const double *d2, int rows, int cols,
                         int spo, double contrast, double prefactor,
                         oracle_extremum *cand, int cand_cap, int *n_cand,
                         oracle_extremum *low, int low_cap, int *n_low);
void oracle_gradient(const double *const *dog, int cols, int s, int m, int n, double g[3]);
void oracle_hessian(const double *const *dog, int cols, int s, int m, int n, double h[3][3]);
int oracle_inverse3x3(const double m[3][3], double inv[3][3]);

void oracle_pyramid_free(oracle_pyramid *p);
int oracle_compute_gaussian_scale_space(const double *input, int rows, int cols,
                                        int number_of_octaves, int scales_per_octave,
                                        double min_blur_level, double assumed_blur,
                                        int separable, oracle_pyramid *p);
int oracle_compute_dog(oracle_pyramid *p);
int oracle_find_candidates(const oracle_pyramid *p, double contrast, double prefactor,
                           oracle_candidate *cand, int cand_cap, int *n_cand,
                           oracle_candidate *low, int low_cap, int *n_low);
int oracle_refine_one(const oracle_pyramid *p, const oracle_candidate *cnd,
                      double contrast, double edge_ratio, int max_iterations, double offset_bound,
                      double min_blur_level, double min_interpixel_distance,
                      oracle_keypoint *kp);
/* test diagnostic: distance of a candidate's refinement walk to the nearest decision flip (see sift_oracle.c) */
double oracle_refine_margin(const oracle_pyramid *p, const oracle_candidate *cnd,
                            double contrast, double edge_ratio, int max_iterations, double offset_bound);
int oracle_refine(const oracle_pyramid *p, const oracle_candidate *cand, int n_cand,
                  double contrast, double edge_ratio, int max_iterations, double offset_bound,
                  double min_blur_level, double min_interpixel_distance,
                  oracle_keypoint *out, int cap, int *n_out, int outcomes[ORACLE_REFINE_NOUTCOMES]);
int oracle_detect(const double *input, int rows, int cols, const oracle_params *prm, int separable,
                  oracle_pyramid *pyr, oracle_keypoint *out, int cap, int *n_out,
                  int *n_cand_out, int *n_low_out, int outcomes[ORACLE_REFINE_NOUTCOMES]);

#ifdef __cplusplus
}
#endif
#endif
