/*
 * sift_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Float64 CPU restatement of the detection path of
 * bingjetli/sift-scale-space-extrema-detection (browser JavaScript).  It is the
 * checker for the CUDA engine: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libsift_b200.so and the host mirror above it) never links or calls it.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures
 *  and no JavaScript engine exists in the build container,
 * so the pin is tests/golden/ref_*.npz: outputs of the UNMODIFIED reference
 * sources (read from /root/reference, never copied) executed by oracle/jsmini.py,
 * a small ECMAScript-subset interpreter, driven like main.js drives the worker
 * (oracle/make_golden.py is the committed generator).  tests/test_golden_reference.py
 * requires this file to reproduce every Gaussian level, DoG level, candidate and
 * refined keypoint of those runs BIT FOR BIT (absoluteSigma to 2 ulp: pow).  The
 * inputs are tiny (the interpreter manages ~60k kernel taps per second); larger
 * sizes rest on this restatement plus the analytic KATs in tests/test_oracle_kat.py.
 *
 * Every function cites the reference file:line it restates.  All arithmetic is
 * IEEE binary64 like JS Number; loop orders and operation orders follow the
 * reference so that results agree to the last few ulps (libm exp/pow may differ
 * from V8's by 1 ulp).  Build with -ffp-contract=off (no FMA contraction: JS
 * never fuses).
 *
 * Images are row-major double[rows*cols]; "Matrix2D m[i][j]" == m[i*cols+j]
 * (matrix2d.js:5-31).
 */
#include <pthread.h>
#include <unistd.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "sift_oracle.h"

#define JS_EPSILON 2.220446049250313e-16 /* Number.EPSILON, matrix2d.js:482 */

/* Math.round: nearest integer, ties toward +infinity.
 * x - floor(x) is exact in binary64, so this is the exact ECMAScript result. */
double oracle_js_round(double x)
{
  double f = floor(x);
  return (x - f >= 0.5) ? f + 1.0 : f;
}

/* ------------------------------------------------------ display products -- */

/* The ImageData the reference posts beside a stage result:
 *   GRAY     image-utils.js:171-220  gray path: Math.round(v * 255) stored into a Uint8ClampedArray
 *            (clamped to 0..255, NaN -> 0), alpha 255
 *   SIGMOID  matrix2d.js:148-156     1 / (1 + Math.exp(coefficient * (-1 * v))) first (background.js:303-307)
 *   MINMAX   matrix2d.js:169-192     (v - min) / (max - min), min/max sampled over the matrix, starting
 *            from MAX_SAFE_INTEGER / MIN_SAFE_INTEGER (background.js:336) */
int oracle_preview(const double *m, int rows, int cols, int mode, double coefficient,
                   unsigned char *rgba, double *min_max)
{
  double mn = 9007199254740991.0, mx = -9007199254740991.0;
  if (!m || !rgba || rows < 1 || cols < 1) return -1;
  if (mode == ORACLE_PREVIEW_MINMAX) {
    for (int i = 0; i < rows; i++)
      for (int j = 0; j < cols; j++) {
        const double v = m[(size_t)i * cols + j];
        if (v < mn) mn = v;
        if (v > mx) mx = v;
      }
  }
  if (min_max) { min_max[0] = mode == ORACLE_PREVIEW_MINMAX ? mn : 0.0; min_max[1] = mode == ORACLE_PREVIEW_MINMAX ? mx : 1.0; }
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      const double v = m[(size_t)i * cols + j];
      double t = v;
      if (mode == ORACLE_PREVIEW_SIGMOID) t = 1 / (1 + exp(coefficient * (-1 * v)));
      else if (mode == ORACLE_PREVIEW_MINMAX) t = (v - mn) / (mx - mn);
      const double r = oracle_js_round(t * 255);
      const unsigned char g = (unsigned char)(r >= 255.0 ? 255 : (r >= 0.0 ? (int)r : 0));   /* NaN -> 0 */
      unsigned char *px = rgba + ((size_t)i * cols + j) * 4;
      px[0] = g; px[1] = g; px[2] = g; px[3] = 255;
    }
  return 0;
}

/* ---------------------------------------------------------------- resize -- */

/* matrix2d.js:112-138  Matrix2D_linearResize.  The loop counters are doubles
 * advanced by sampling_rate; samples are in[floor(i)][floor(j)]. */
int oracle_linear_resize_dims(int rows, int cols, double rate, int *orows, int *ocols)
{
  int r = 0, c = 0;
  for (double i = 0; i < rows; i += rate) r++;   /* matrix2d.js:119 */
  for (double j = 0; j < cols; j += rate) c++;   /* matrix2d.js:124 */
  *orows = r; *ocols = c;
  return 0;
}

int oracle_linear_resize(const double *in, int rows, int cols, double rate, double *out)
{
  int orows, ocols;
  oracle_linear_resize_dims(rows, cols, rate, &orows, &ocols);
  int a = 0;
  for (double i = 0; i < rows; i += rate, a++) {
    int b = 0;
    for (double j = 0; j < cols; j += rate, b++) {
      /* Number(x.toString()) round-trips a double exactly (matrix2d.js:129) */
      out[(size_t)a * ocols + b] = in[(size_t)((int)floor(i)) * cols + (int)floor(j)];
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- kernel -- */

/* sift.js:22-26  sample2DGaussian */
static double sample_2d_gaussian(double i, double j, double sigma)
{
  return exp((((i * i) + (j * j)) / (sigma * sigma)) * -0.5) / (2 * M_PI * (sigma * sigma));
}

/* sift.js:38,44 */
int oracle_kernel_radius(double sigma)
{
  return (int)oracle_js_round(3 * sigma);
}

/* sift.js:31-67  buildGaussianKernel.  kernel[i][j], i outer, j inner; the
 * normaliser is the running sum in that order. `kernel` holds size*size doubles
 * where size = 2*round(3 sigma)+1. */
int oracle_build_gaussian_kernel(double sigma, double *kernel)
{
  const int offset = oracle_kernel_radius(sigma);
  const int size = 2 * offset + 1;
  double sum = 0;
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++) {
      double g = sample_2d_gaussian(i - offset, j - offset, sigma);
      sum += g;
      kernel[(size_t)i * size + j] = g;
    }
  for (int i = 0; i < size; i++)
    for (int j = 0; j < size; j++)
      kernel[(size_t)i * size + j] = kernel[(size_t)i * size + j] / sum;
  return size;
}

/* ------------------------------------------------------------------ blur -- */

/* sift.js:72-149  SIFT_blurMatrix2DChunk.  Dense 2D correlation, sample
 * coordinates clamped per axis, i = x offset (outer), j = y offset (inner),
 * separate multiply and add. Writes output[y][x] for the half-open chunk. */
int oracle_blur_chunk(const double *input, int rows, int cols, double *output,
                      double sigma, int x1, int y1, int x2, int y2)
{
  const int R = oracle_kernel_radius(sigma);
  const int size = 2 * R + 1;
  double *kernel = (double *)malloc(sizeof(double) * (size_t)size * size);
  if (!kernel) return -1;
  oracle_build_gaussian_kernel(sigma, kernel);     /* rebuilt per chunk, sift.js:81 */
  const int shift = size / 2;                      /* sift.js:83 */
  for (int y = y1; y < y2; y++) {
    for (int x = x1; x < x2; x++) {
      double pixel_sum = 0;
      for (int i = 0; i < size; i++) {
        int _x = x + (i - shift);
        if (_x < 0) _x = 0; else if (_x >= cols) _x = cols - 1;
        const double *krow = kernel + (size_t)i * size;
        for (int j = 0; j < size; j++) {
          int _y = y + (j - shift);
          if (_y < 0) _y = 0; else if (_y >= rows) _y = rows - 1;
          pixel_sum += input[(size_t)_y * cols + _x] * krow[j];   /* sift.js:125 */
        }
      }
      output[(size_t)y * cols + x] = pixel_sum;
    }
  }
  free(kernel);
  return 0;
}

/* Whole-image blur = union of the 32x32 chunks of background.js:181-203;
 * every output pixel is independent of the chunking (sift.js:137). */
int oracle_blur_image(const double *input, int rows, int cols, double *output, double sigma)
{
  return oracle_blur_chunk(input, rows, cols, output, sigma, 0, 0, cols, rows);
}

/* Separable float64 variant for sizes where the dense kernel is too slow.
 * NOT line-by-line: mathematically the 2D kernel is the outer product of the
 * normalised 1D kernel (g(i,j) = g1(i) g1(j), sum = (sum g1)^2); results agree
 * with oracle_blur_image to ~1e-15 relative (tests/test_oracle.py checks). */
/* Row bands of the two separable passes.  Every output is an independent sum taken in ascending tap order, so
 * bands can run on separate threads (pthreads; ORACLE_THREADS, default = online cores, at most 32) and the
 * vertical pass can run tap-outer / column-inner (contiguous reads) without changing a single bit of the result. */
typedef struct {
  const double *in; double *out; const double *k1;
  int rows, cols, R, y0, y1, vertical;
} sep_band;

static void *sep_band_run(void *arg)
{
  const sep_band *b = (const sep_band *)arg;
  const int size = 2 * b->R + 1, R = b->R, rows = b->rows, cols = b->cols;
  if (!b->vertical) {
    for (int y = b->y0; y < b->y1; y++)
      for (int x = 0; x < cols; x++) {
        double acc = 0;
        for (int i = 0; i < size; i++) {
          int _x = x + i - R; if (_x < 0) _x = 0; else if (_x >= cols) _x = cols - 1;
          acc += b->in[(size_t)y * cols + _x] * b->k1[i];
        }
        b->out[(size_t)y * cols + x] = acc;
      }
  } else {
    for (int y = b->y0; y < b->y1; y++) {
      double *out = b->out + (size_t)y * cols;
      for (int x = 0; x < cols; x++) out[x] = 0;
      for (int j = 0; j < size; j++) {
        int _y = y + j - R; if (_y < 0) _y = 0; else if (_y >= rows) _y = rows - 1;
        const double *t = b->in + (size_t)_y * cols;
        const double w = b->k1[j];
        for (int x = 0; x < cols; x++) out[x] += t[x] * w;
      }
    }
  }
  return NULL;
}

static int sep_threads(int rows)
{
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  const char *e = getenv("ORACLE_THREADS");
  if (e && atoi(e) > 0) n = atoi(e);
  if (n > 32) n = 32;
  if (n > rows / 16) n = rows / 16;
  return n < 1 ? 1 : (int)n;
}

static void sep_pass(const double *in, double *out, const double *k1, int rows, int cols, int R, int vertical)
{
  const int nt = sep_threads(rows);
  pthread_t th[32];
  sep_band band[32];
  for (int t = 0; t < nt; t++) {
    sep_band b = { in, out, k1, rows, cols, R, (int)((long long)rows * t / nt), (int)((long long)rows * (t + 1) / nt), vertical };
    band[t] = b;
  }
  int started = 0;
  for (int t = 1; t < nt; t++) {
    if (pthread_create(&th[t], NULL, sep_band_run, &band[t]) != 0) break;
    started = t;
  }
  sep_band_run(&band[0]);
  for (int t = started + 1; t < nt; t++) sep_band_run(&band[t]);      /* threads that could not be created */
  for (int t = 1; t <= started; t++) pthread_join(th[t], NULL);
}

int oracle_blur_image_separable(const double *input, int rows, int cols, double *output, double sigma)
{
  const int R = oracle_kernel_radius(sigma);
  const int size = 2 * R + 1;
  double *k1 = (double *)malloc(sizeof(double) * size);
  double *tmp = (double *)malloc(sizeof(double) * (size_t)rows * cols);
  if (!k1 || !tmp) { free(k1); free(tmp); return -1; }
  double s = 0;
  for (int i = 0; i < size; i++) { k1[i] = exp(-0.5 * (double)((i - R) * (i - R)) / (sigma * sigma)); s += k1[i]; }
  for (int i = 0; i < size; i++) k1[i] /= s;
  sep_pass(input, tmp, k1, rows, cols, R, 0);
  sep_pass(tmp, output, k1, rows, cols, R, 1);
  free(k1); free(tmp);
  return 0;
}

/* sift.js:154-188  SIFT_subtractMatrix2DChunk: pair[0] - pair[1]. */
int oracle_subtract_chunk(const double *a, const double *b, int rows, int cols, double *output,
                          int x1, int y1, int x2, int y2)
{
  (void)rows;
  for (int y = y1; y < y2; y++)
    for (int x = x1; x < x2; x++)
      output[(size_t)y * cols + x] = a[(size_t)y * cols + x] - b[(size_t)y * cols + x];
  return 0;
}

/* --------------------------------------------------------------- extrema -- */

/* sift.js:285 (and background.js:572): ((2^(1/spo)-1)/(2^(1/3)-1)) * C */
double oracle_contrast_threshold(int spo, double contrast)
{
  return ((pow(2, 1.0 / spo) - 1) / (pow(2, 1.0 / 3) - 1)) * contrast;
}

/* sift.js:212-316  SIFT_findExtremas.  Strict 26-neighbour min or max, raster
 * order, then abs(c) >= threshold*prefactor splits candidates / low contrast.
 * Returns the number of extrema of the requested class written (at most cap)
 * and stores the true counts in n_cand / n_low. */
int oracle_find_extremas(const double *d0, const double *d1, const double *d2, int rows, int cols,
                         int spo, double contrast, double prefactor,
                         oracle_extremum *cand, int cand_cap, int *n_cand,
                         oracle_extremum *low, int low_cap, int *n_low)
{
  const double *t[3] = { d0, d1, d2 };
  int nc = 0, nl = 0;
  const double threshold = oracle_contrast_threshold(spo, contrast);
  const double pixel_threshold = threshold * prefactor;            /* sift.js:293 */
  for (int y = 1; y < rows - 1; y++) {
    for (int x = 1; x < cols - 1; x++) {
      const double c = d1[(size_t)y * cols + x];
      int is_min = 1, is_max = 1;
      for (int p = 0; p < 3; p++)
        for (int dy = -1; dy <= 1; dy++)
          for (int dx = -1; dx <= 1; dx++) {
            if (p == 1 && dy == 0 && dx == 0) continue;
            double v = t[p][(size_t)(y + dy) * cols + (x + dx)];
            if (!(v > c)) is_min = 0;                               /* sift.js:261 */
            if (!(v < c)) is_max = 0;                               /* sift.js:266 */
          }
      if (is_min || is_max) {
        if (fabs(c) >= pixel_threshold) {                           /* sift.js:294 */
          if (cand && nc < cand_cap) { cand[nc].x = x; cand[nc].y = y; cand[nc].value = c; }
          nc++;
        } else {
          if (low && nl < low_cap) { low[nl].x = x; low[nl].y = y; low[nl].value = c; }
          nl++;
        }
      }
    }
  }
  if (n_cand) *n_cand = nc;
  if (n_low) *n_low = nl;
  return 0;
}

/* ------------------------------------------------- gradient / Hessian / 3x3 */

#define DOG(s_, m_, n_) (dog[(s_)][(size_t)(m_) * cols + (n_)])

/* sift.js:333-353  order [s, m(row), n(col)] */
void oracle_gradient(const double *const *dog, int cols, int s, int m, int n, double g[3])
{
  g[0] = (DOG(s + 1, m, n) - DOG(s - 1, m, n)) / 2;
  g[1] = (DOG(s, m + 1, n) - DOG(s, m - 1, n)) / 2;
  g[2] = (DOG(s, m, n + 1) - DOG(s, m, n - 1)) / 2;
}

/* sift.js:377-447 */
void oracle_hessian(const double *const *dog, int cols, int s, int m, int n, double h[3][3])
{
  const double h11 = (DOG(s + 1, m, n) + DOG(s - 1, m, n) - (2 * DOG(s, m, n)));
  const double h22 = (DOG(s, m + 1, n) + DOG(s, m - 1, n) - (2 * DOG(s, m, n)));
  const double h33 = (DOG(s, m, n + 1) + DOG(s, m, n - 1) - (2 * DOG(s, m, n)));
  const double h12 = (DOG(s + 1, m + 1, n) - DOG(s + 1, m - 1, n) - DOG(s - 1, m + 1, n) + DOG(s - 1, m - 1, n)) / 4;
  const double h13 = (DOG(s + 1, m, n + 1) - DOG(s + 1, m, n - 1) - DOG(s - 1, m, n + 1) + DOG(s - 1, m, n - 1)) / 4;
  const double h23 = (DOG(s, m + 1, n + 1) - DOG(s, m + 1, n - 1) - DOG(s, m - 1, n + 1) + DOG(s, m - 1, n - 1)) / 4;
  h[0][0] = h11; h[0][1] = h12; h[0][2] = h13;
  h[1][0] = h12; h[1][1] = h22; h[1][2] = h23;
  h[2][0] = h13; h[2][1] = h23; h[2][2] = h33;
}
#undef DOG

/* matrix2d.js:349-382 + 197-212: determinant of the matrix with row i and
 * column j removed: (a*d) - (b*c) over the remaining 2x2 in row-major order. */
static double minor2x2(const double m[3][3], int i, int j)
{
  double v[4]; int k = 0;
  for (int r = 0; r < 3; r++) {
    if (r == i) continue;
    for (int c = 0; c < 3; c++) {
      if (c == j) continue;
      v[k++] = m[r][c];
    }
  }
  return (v[0] * v[3]) - (v[1] * v[2]);
}

/* matrix2d.js:464-509  Matrix2D_get3x3Inverse: returns 0 ("null") when
 * abs(det) < Number.EPSILON, else 1 with inv = transpose(cofactors) / det. */
int oracle_inverse3x3(const double m[3][3], double inv[3][3])
{
  double minors[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      minors[i][j] = minor2x2(m, i, j);                      /* matrix2d.js:258-262, 303-336 */
  const double det = (m[0][0] * minors[0][0]) - (m[0][1] * minors[0][1]) + (m[0][2] * minors[0][2]); /* 265-269 */
  if (fabs(det) < JS_EPSILON) return 0;                      /* matrix2d.js:482 */
  double cof[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      cof[i][j] = minors[i][j] * (((i + j) & 1) ? -1.0 : 1.0); /* matrix2d.js:406, Math.pow(-1,i+j) */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      inv[i][j] = cof[j][i] / det;                           /* transpose 438, scalarDivide 448 */
  return 1;
}

/* ------------------------------------------------------------- pipeline -- */

void oracle_pyramid_free(oracle_pyramid *p)
{
  if (!p) return;
  for (int o = 0; o < ORACLE_MAX_OCTAVES; o++)
    for (int s = 0; s < ORACLE_MAX_LEVELS; s++) {
      free(p->gauss[o][s]); p->gauss[o][s] = NULL;
      free(p->dog[o][s]); p->dog[o][s] = NULL;
    }
}

/* background.js:71-237  computeGaussianScaleSpace.  separable != 0 swaps the
 * dense blur for oracle_blur_image_separable (large sizes only). */
int oracle_compute_gaussian_scale_space(const double *input, int rows, int cols,
                                        int number_of_octaves, int scales_per_octave,
                                        double min_blur_level, double assumed_blur,
                                        int separable, oracle_pyramid *p)
{
  if (number_of_octaves > ORACLE_MAX_OCTAVES || scales_per_octave + 3 > ORACLE_MAX_LEVELS) return -2;
  memset(p, 0, sizeof(*p));
  p->octaves = number_of_octaves;
  p->levels = scales_per_octave + 3;
  p->spo = scales_per_octave;

  int brows, bcols;
  oracle_linear_resize_dims(rows, cols, 0.5, &brows, &bcols);
  double *base_image = (double *)malloc(sizeof(double) * (size_t)brows * bcols);
  if (!base_image) return -1;
  oracle_linear_resize(input, rows, cols, 0.5, base_image);          /* background.js:84 */
  int base_owned = 1;                                                /* octave 0 base is not a level */
  double base_blur_level = min_blur_level;                           /* :89 */
  const double k = pow(2, 1.0 / scales_per_octave);                  /* :100 */

  for (int octave = 0; octave < number_of_octaves; octave++) {
    for (int scale = 0; scale < scales_per_octave + 3; scale++) {
      if (octave > 0 && scale == 0) {
        const int prow = p->rows[octave - 1], pcol = p->cols[octave - 1];
        const double *seed = p->gauss[octave - 1][scales_per_octave]; /* :114 */
        if (base_owned) { free(base_image); base_owned = 0; }
        oracle_linear_resize_dims(prow, pcol, 2.0, &brows, &bcols);
        base_image = (double *)malloc(sizeof(double) * (size_t)brows * bcols);
        if (!base_image) return -1;
        oracle_linear_resize(seed, prow, pcol, 2.0, base_image);     /* :118 */
        base_blur_level = p->blur[octave - 1][scales_per_octave];    /* :122 */
        p->gauss[octave][0] = base_image;                            /* stored unblurred, :127-130 */
        p->blur[octave][0] = base_blur_level;
      } else {
        double *output = (double *)calloc((size_t)brows * bcols, sizeof(double));
        if (!output) return -1;
        const double current_k = pow(k, scale);                      /* :157 */
        const double target_sigma = base_blur_level * current_k;     /* :173 */
        const double base_sigma = (octave == 0) ? assumed_blur : base_blur_level; /* :174-176 */
        const double offset_sigma = sqrt((target_sigma * target_sigma) - (base_sigma * base_sigma)); /* :177 */
        int rc = separable ? oracle_blur_image_separable(base_image, brows, bcols, output, offset_sigma)
                           : oracle_blur_image(base_image, brows, bcols, output, offset_sigma);
        if (rc) return rc;
        p->gauss[octave][scale] = output;
        p->blur[octave][scale] = target_sigma;                       /* :208 */
        p->offset_sigma[octave][scale] = offset_sigma;
      }
    }
    p->rows[octave] = brows; p->cols[octave] = bcols;
  }
  if (base_owned) free(base_image);
  return 0;
}

/* background.js:258-354  computeDifferenceOfGaussians: D[o][s-1] = S[o][s-1] - S[o][s],
 * blurLevel of D = blurLevel of S[s-1] (:327). */
int oracle_compute_dog(oracle_pyramid *p)
{
  for (int o = 0; o < p->octaves; o++) {
    const size_t n = (size_t)p->rows[o] * p->cols[o];
    for (int scale = 1; scale < p->levels; scale++) {
      double *out = (double *)malloc(sizeof(double) * n);
      if (!out) return -1;
      oracle_subtract_chunk(p->gauss[o][scale - 1], p->gauss[o][scale], p->rows[o], p->cols[o], out,
                            0, 0, p->cols[o], p->rows[o]);
      p->dog[o][scale - 1] = out;
      p->dog_blur[o][scale - 1] = p->blur[o][scale - 1];
    }
  }
  return 0;
}

/* background.js:359-450  findCandidateKeypoints: octave-major, DoG scales
 * 1..nDoG-2, raster order inside a scale. */
int oracle_find_candidates(const oracle_pyramid *p, double contrast, double prefactor,
                           oracle_candidate *cand, int cand_cap, int *n_cand,
                           oracle_candidate *low, int low_cap, int *n_low)
{
  int nc = 0, nl = 0;
  const int ndog = p->levels - 1;
  for (int o = 0; o < p->octaves; o++) {
    for (int scale = 1; scale < ndog - 1; scale++) {
      int c = 0, l = 0;
      const int want = p->rows[o] * p->cols[o];
      oracle_extremum *ec = (oracle_extremum *)malloc(sizeof(oracle_extremum) * (size_t)(want > 0 ? want : 1));
      oracle_extremum *el = (oracle_extremum *)malloc(sizeof(oracle_extremum) * (size_t)(want > 0 ? want : 1));
      if (!ec || !el) { free(ec); free(el); return -1; }
      oracle_find_extremas(p->dog[o][scale - 1], p->dog[o][scale], p->dog[o][scale + 1],
                           p->rows[o], p->cols[o], p->spo, contrast, prefactor,
                           ec, want, &c, el, want, &l);
      for (int i = 0; i < c; i++, nc++)
        if (cand && nc < cand_cap) { cand[nc].octave = o; cand[nc].scale = scale; cand[nc].x = ec[i].x; cand[nc].y = ec[i].y; cand[nc].value = ec[i].value; }
      for (int i = 0; i < l; i++, nl++)
        if (low && nl < low_cap) { low[nl].octave = o; low[nl].scale = scale; low[nl].x = el[i].x; low[nl].y = el[i].y; low[nl].value = el[i].value; }
      free(ec); free(el);
    }
  }
  if (n_cand) *n_cand = nc;
  if (n_low) *n_low = nl;
  return 0;
}

/* background.js:455-685  refineCandidateKeypoints, one candidate.
 * Returns an ORACLE_REFINE_* outcome; fills *kp when accepted.  A singular
 * Hessian (reference: uncaught TypeError) is reported as
 * ORACLE_REFINE_SINGULAR and the candidate is discarded. */
int oracle_refine_one(const oracle_pyramid *p, const oracle_candidate *cnd,
                      double contrast, double edge_ratio, int max_iterations, double offset_bound,
                      double min_blur_level, double min_interpixel_distance,
                      oracle_keypoint *kp)
{
  const int octave = cnd->octave;
  const int rows = p->rows[octave], cols = p->cols[octave];
  const int ndog = p->levels - 1;
  const double *const *dog = (const double *const *)p->dog[octave];
  int s = cnd->scale, m = cnd->y, n = cnd->x;
  for (int i = 0; i < max_iterations; i++) {                        /* :480 */
    double g[3], h[3][3], inv[3][3];
    oracle_gradient(dog, cols, s, m, n, g);                         /* :505 */
    oracle_hessian(dog, cols, s, m, n, h);                          /* :523 */
    if (!oracle_inverse3x3(h, inv)) return ORACLE_REFINE_SINGULAR;  /* :546 -> TypeError at :552 */
    double alpha[3];
    for (int r = 0; r < 3; r++) {                                   /* :551-554, matrix2d.js:454, 514-541 */
      double result = 0;
      for (int c = 0; c < 3; c++) result += (inv[r][c] * -1) * g[c];
      alpha[r] = result;
    }
    if (fabs(alpha[0]) < offset_bound && fabs(alpha[1]) < offset_bound && fabs(alpha[2]) < offset_bound) { /* :558 */
      const double interpolated_value = cnd->value +
          (((0.5 * alpha[0]) * g[0]) + ((0.5 * alpha[1]) * g[1]) + ((0.5 * alpha[2]) * g[2])); /* :565, Q5 */
      const double threshold = oracle_contrast_threshold(p->spo, contrast);                    /* :572 */
      if (fabs(interpolated_value) < threshold) return ORACLE_REFINE_LOW_CONTRAST;             /* :577 */
      const double tr = 0 + h[1][1] + h[2][2];                      /* :589-592, matrix2d.js:555-558 */
      const double det = (h[1][1] * h[2][2]) - (h[1][2] * h[2][1]); /* :593 */
      const double edgeness = (tr * tr) / det;                      /* :594 */
      const double edge_threshold = ((edge_ratio + 1) * (edge_ratio + 1)) / edge_ratio;        /* :598 */
      if (edgeness > edge_threshold) return ORACLE_REFINE_EDGE;     /* :599, Q6: NaN and negatives pass */
      const double delta = pow(2, octave - 1);                      /* :611 */
      kp->octave = octave;
      kp->scaleLevel = s; kp->localX = n; kp->localY = m;           /* :620-623 */
      kp->absoluteY = delta * (alpha[1] + m);                       /* :612 */
      kp->absoluteX = delta * (alpha[2] + n);                       /* :613 */
      kp->absoluteSigma = (delta / min_interpixel_distance) * min_blur_level *
                          pow(2, (alpha[0] + s) / p->spo);          /* :614 */
      kp->interpolatedValue = interpolated_value;
      kp->offset[0] = alpha[0]; kp->offset[1] = alpha[1]; kp->offset[2] = alpha[2];
      kp->dogValue = cnd->value;
      kp->candScale = cnd->scale; kp->candX = cnd->x; kp->candY = cnd->y;
      kp->iterations = i;
      return ORACLE_REFINE_ACCEPTED;
    }
    s = (int)oracle_js_round(s + alpha[0]);                         /* :638-640 */
    m = (int)oracle_js_round(m + alpha[1]);
    n = (int)oracle_js_round(n + alpha[2]);
    if (s < 1 || s >= ndog - 1) return ORACLE_REFINE_LEFT_SCALE;    /* :644 */
    if (m < 1 || m >= rows - 1) return ORACLE_REFINE_LEFT_ROWS;     /* :651 */
    if (n < 1 || n >= cols - 1) return ORACLE_REFINE_LEFT_COLS;     /* :658 */
  }
  return ORACLE_REFINE_NO_CONVERGENCE;
}

/* TEST DIAGNOSTIC (no reference counterpart): walks a candidate exactly as oracle_refine_one does and returns how
 * close the walk came to flipping a decision -- the smallest relative distance to a comparison threshold it met
 * (offset bound :558, contrast :577, edge :599) or absolute distance of a moved coordinate to a Math.round boundary
 * (:638-640).  A mismatch between two implementations is "explained" when this margin is within the stated
 * tolerance (north_star: 1e-5 of a threshold). */
double oracle_refine_margin(const oracle_pyramid *p, const oracle_candidate *cnd,
                            double contrast, double edge_ratio, int max_iterations, double offset_bound)
{
  const int octave = cnd->octave;
  const int rows = p->rows[octave], cols = p->cols[octave];
  const int ndog = p->levels - 1;
  const double *const *dog = (const double *const *)p->dog[octave];
  int s = cnd->scale, m = cnd->y, n = cnd->x;
  double margin = INFINITY;
#define MARGIN_(d) do { const double d_ = (d); if (d_ < margin) margin = d_; } while (0)
  for (int i = 0; i < max_iterations; i++) {
    double g[3], h[3][3], inv[3][3];
    oracle_gradient(dog, cols, s, m, n, g);
    oracle_hessian(dog, cols, s, m, n, h);
    if (!oracle_inverse3x3(h, inv)) return 0.0;
    double alpha[3];
    for (int r = 0; r < 3; r++) {
      double result = 0;
      for (int c = 0; c < 3; c++) result += (inv[r][c] * -1) * g[c];
      alpha[r] = result;
    }
    for (int r = 0; r < 3; r++) MARGIN_(fabs(fabs(alpha[r]) - offset_bound) / offset_bound);
    if (fabs(alpha[0]) < offset_bound && fabs(alpha[1]) < offset_bound && fabs(alpha[2]) < offset_bound) {
      const double v = cnd->value + (((0.5 * alpha[0]) * g[0]) + ((0.5 * alpha[1]) * g[1]) + ((0.5 * alpha[2]) * g[2]));
      const double threshold = oracle_contrast_threshold(p->spo, contrast);
      MARGIN_(fabs(fabs(v) - threshold) / threshold);
      if (fabs(v) < threshold) return margin;
      const double tr = 0 + h[1][1] + h[2][2];
      const double det = (h[1][1] * h[2][2]) - (h[1][2] * h[2][1]);
      const double edgeness = (tr * tr) / det;
      const double edge_threshold = ((edge_ratio + 1) * (edge_ratio + 1)) / edge_ratio;
      if (edgeness == edgeness) MARGIN_(fabs(edgeness - edge_threshold) / edge_threshold);
      return margin;
    }
    const double moved[3] = { s + alpha[0], m + alpha[1], n + alpha[2] };
    for (int r = 0; r < 3; r++) {
      const double f = (moved[r] + 0.5) - floor(moved[r] + 0.5);      /* 0 on a rounding boundary */
      MARGIN_(f < 1 - f ? f : 1 - f);
    }
    s = (int)oracle_js_round(moved[0]); m = (int)oracle_js_round(moved[1]); n = (int)oracle_js_round(moved[2]);
    if (s < 1 || s >= ndog - 1 || m < 1 || m >= rows - 1 || n < 1 || n >= cols - 1) return margin;
  }
#undef MARGIN_
  return margin;
}

/* background.js:455-685 over the candidate list in reference order. */
int oracle_refine(const oracle_pyramid *p, const oracle_candidate *cand, int n_cand,
                  double contrast, double edge_ratio, int max_iterations, double offset_bound,
                  double min_blur_level, double min_interpixel_distance,
                  oracle_keypoint *out, int cap, int *n_out, int outcomes[ORACLE_REFINE_NOUTCOMES])
{
  int n = 0;
  if (outcomes) memset(outcomes, 0, sizeof(int) * ORACLE_REFINE_NOUTCOMES);
  for (int i = 0; i < n_cand; i++) {
    oracle_keypoint kp;
    int rc = oracle_refine_one(p, &cand[i], contrast, edge_ratio, max_iterations, offset_bound,
                               min_blur_level, min_interpixel_distance, &kp);
    if (outcomes) outcomes[rc]++;
    if (rc == ORACLE_REFINE_ACCEPTED) {
      if (out && n < cap) out[n] = kp;
      n++;
    }
  }
  if (n_out) *n_out = n;
  return 0;
}

/* Whole detection path in the order main.js:111 -> 239 -> 274 -> 325 drives it. */
int oracle_detect(const double *input, int rows, int cols, const oracle_params *prm, int separable,
                  oracle_pyramid *pyr, oracle_keypoint *out, int cap, int *n_out,
                  int *n_cand_out, int *n_low_out, int outcomes[ORACLE_REFINE_NOUTCOMES])
{
  int rc = oracle_compute_gaussian_scale_space(input, rows, cols, prm->numberOfOctaves, prm->scalesPerOctave,
                                               prm->minBlurLevel, prm->assumedBlur, separable, pyr);
  if (rc) return rc;
  rc = oracle_compute_dog(pyr);
  if (rc) return rc;
  int nc = 0, nl = 0;
  oracle_find_candidates(pyr, prm->contrastThreshold, prm->preFilterFactor, NULL, 0, &nc, NULL, 0, &nl);
  oracle_candidate *cand = (oracle_candidate *)malloc(sizeof(oracle_candidate) * (size_t)(nc > 0 ? nc : 1));
  if (!cand) return -1;
  oracle_find_candidates(pyr, prm->contrastThreshold, prm->preFilterFactor, cand, nc, &nc, NULL, 0, &nl);
  rc = oracle_refine(pyr, cand, nc, prm->contrastThreshold, prm->edgeRatio, prm->maxIterations, prm->offsetBound,
                     prm->minBlurLevel, prm->minInterpixelDistance, out, cap, n_out, outcomes);
  free(cand);
  if (n_cand_out) *n_cand_out = nc;
  if (n_low_out) *n_low_out = nl;
  return rc;
}
