// refine.cu -- quadratic sub-pixel refinement, one thread per candidate.
// COMPILED WITH --fmad=false: the reference is JavaScript float64 and never fuses a
// multiply with an add, so the operation order below reproduces its rounding given
// the same DoG samples.
//
// Restates refineCandidateKeypoints (background.js:455-685): up to maxIterations
// rounds of  g = SIFT_generateGradientVector (src/sift.js:333-353),
// H = SIFT_generateHessianMatrix (src/sift.js:377-447), alpha = (-H^-1) g with the
// cofactor inverse Matrix2D_get3x3Inverse (src/matrix2d.js:464-509, abs(det) <
// Number.EPSILON -> no inverse), acceptance when all abs(alpha) < 0.6, contrast test
// on extrema.value + 0.5 alpha.g (background.js:565-583, uses the ORIGINAL
// candidate value), edge test on the spatial 2x2 block (background.js:589-604, no
// det <= 0 rejection), absolute coordinates (background.js:611-614); otherwise the
// sample moves by Math.round (ties up) and is dropped when it leaves the DoG
// volume (background.js:638-664).
#include "common.cuh"

#define JS_EPSILON 2.220446049250313e-16

__device__ __forceinline__ double js_round(double x)
{
  const double f = floor(x);
  return (x - f >= 0.5) ? f + 1.0 : f;
}

template <typename T> struct DogView {
  const T *pl[SIFT_MAX_LEVELS];
  size_t pitch;
  __device__ __forceinline__ double at(int s, int m, int n) const { return (double)pl[s][(size_t)m * pitch + n]; }
};

template <typename V>
__device__ __forceinline__ void grad_hess(const V &D, int s, int m, int n, double g[3], double h[3][3])
{
  const double c = D.at(s, m, n);
  const double sp = D.at(s + 1, m, n), sm = D.at(s - 1, m, n);
  const double mp = D.at(s, m + 1, n), mm = D.at(s, m - 1, n);
  const double np = D.at(s, m, n + 1), nm = D.at(s, m, n - 1);
  g[0] = (sp - sm) / 2;                                          // sift.js:336-339
  g[1] = (mp - mm) / 2;                                          // sift.js:342-345
  g[2] = (np - nm) / 2;                                          // sift.js:348-351
  const double h11 = (sp + sm - (2 * c));                        // sift.js:380-384
  const double h22 = (mp + mm - (2 * c));                        // sift.js:386-390
  const double h33 = (np + nm - (2 * c));                        // sift.js:392-396
  const double h12 = (D.at(s + 1, m + 1, n) - D.at(s + 1, m - 1, n) - D.at(s - 1, m + 1, n) + D.at(s - 1, m - 1, n)) / 4;
  const double h13 = (D.at(s + 1, m, n + 1) - D.at(s + 1, m, n - 1) - D.at(s - 1, m, n + 1) + D.at(s - 1, m, n - 1)) / 4;
  const double h23 = (D.at(s, m + 1, n + 1) - D.at(s, m + 1, n - 1) - D.at(s, m - 1, n + 1) + D.at(s, m - 1, n - 1)) / 4;
  h[0][0] = h11; h[0][1] = h12; h[0][2] = h13;
  h[1][0] = h12; h[1][1] = h22; h[1][2] = h23;
  h[2][0] = h13; h[2][1] = h23; h[2][2] = h33;
}

// matrix2d.js:349-382 + 197-212
__device__ __forceinline__ double minor2x2(const double m[3][3], int i, int j)
{
  const int r0 = (i == 0) ? 1 : 0, r1 = (i == 2) ? 1 : 2;
  const int c0 = (j == 0) ? 1 : 0, c1 = (j == 2) ? 1 : 2;
  return (m[r0][c0] * m[r1][c1]) - (m[r0][c1] * m[r1][c0]);
}

// matrix2d.js:464-509; returns false for the reference's `null`.
__device__ __forceinline__ bool inverse3x3(const double m[3][3], double inv[3][3])
{
  double minors[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) minors[i][j] = minor2x2(m, i, j);
  const double det = (m[0][0] * minors[0][0]) - (m[0][1] * minors[0][1]) + (m[0][2] * minors[0][2]);
  if (fabs(det) < JS_EPSILON) return false;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double cof = minors[j][i] * (((i + j) & 1) ? -1.0 : 1.0);   // cofactor (j,i): transpose
      inv[i][j] = cof / det;
    }
  return true;
}

// One walk of the loop at background.js:480-668, from any iteration (a fresh candidate starts at it = 0 on its
// own pixel; a walk handed over by another strip resumes where it left).  m is a row of the whole image; the
// strip holds rows [ytop, ytop + oc.h).
__device__ __forceinline__ void refine_walk(const OctaveDev *__restrict__ octs, const sift_walk w0, const RefineParams &rp,
                                            sift_keypoint *__restrict__ out, int cap, Counters *ctr,
                                            sift_walk *__restrict__ walks, int walk_cap)
{
  const OctaveDev &oc = octs[w0.octave];
  DogView<float> D;
#pragma unroll
  for (int i = 0; i < SIFT_MAX_LEVELS; i++) D.pl[i] = oc.dog[i];
  D.pitch = oc.pitch;
  const int rows = oc.gh, cols = oc.w, ytop = oc.y_top;
  int s = w0.scaleLevel, m = w0.y, nn = w0.x, it = w0.iteration;
  const double value = (double)w0.value;
  int outcome = REFINE_NO_CONVERGENCE;
  // the 3x3x3 neighbourhood of the sample must lie in rows that hold the whole image's values
  bool left_strip = m - ytop - 1 < oc.valid0 || m - ytop + 1 >= oc.valid1;     // (handed to the wrong strip)
  sift_keypoint kp;
  for (; !left_strip && it < rp.max_iter; it++) {                                // background.js:480
    double g[3], h[3][3], inv[3][3];
    grad_hess(D, s, m - ytop, nn, g, h);
    if (!inverse3x3(h, inv)) { outcome = REFINE_SINGULAR; break; }               // matrix2d.js:482 (Q7)
    double a[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {                                                // matrix2d.js:454, 527-537
      double res = 0;
#pragma unroll
      for (int c = 0; c < 3; c++) res += (inv[r][c] * -1) * g[c];
      a[r] = res;
    }
    if (fabs(a[0]) < rp.offset_bound && fabs(a[1]) < rp.offset_bound && fabs(a[2]) < rp.offset_bound) {  // :558
      const double omega = value + (((0.5 * a[0]) * g[0]) + ((0.5 * a[1]) * g[1]) + ((0.5 * a[2]) * g[2])); // :565
      if (fabs(omega) < rp.contrast_thr) { outcome = REFINE_LOW_CONTRAST; break; }                          // :577
      const double tr = 0 + h[1][1] + h[2][2];                                   // :592
      const double det = (h[1][1] * h[2][2]) - (h[1][2] * h[2][1]);              // :593
      const double edgeness = (tr * tr) / det;                                   // :594
      if (edgeness > rp.edge_thr) { outcome = REFINE_EDGE; break; }              // :599 (NaN / negative pass, Q6)
      const double delta = exp2((double)(w0.octave - 1));                        // :611 Math.pow(2, octave-1), exact
      kp.octave = w0.octave; kp.scaleLevel = s; kp.localX = nn; kp.localY = m;
      kp.absoluteY = delta * (a[1] + m);                                         // :612
      kp.absoluteX = delta * (a[2] + nn);                                        // :613
      kp.absoluteSigma = (delta / rp.min_interpixel) * rp.min_blur * pow(2.0, (a[0] + s) / rp.spo);         // :614
      kp.interpolatedValue = omega;
      kp.offset[0] = (float)a[0]; kp.offset[1] = (float)a[1]; kp.offset[2] = (float)a[2];
      kp.dogValue = w0.value;
      kp.candScale = w0.candScale; kp.candX = w0.candX; kp.candY = w0.candY; kp.iterations = it;
      outcome = REFINE_ACCEPTED;
      break;
    }
    s = (int)js_round(s + a[0]);                                                 // :638-640
    m = (int)js_round(m + a[1]);
    nn = (int)js_round(nn + a[2]);
    if (s < 1 || s >= rp.ndog - 1) { outcome = REFINE_LEFT_SCALE; break; }       // :644
    if (m < 1 || m >= rows - 1) { outcome = REFINE_LEFT_ROWS; break; }           // :651
    if (nn < 1 || nn >= cols - 1) { outcome = REFINE_LEFT_COLS; break; }         // :658
    if (m - ytop - 1 < oc.valid0 || m - ytop + 1 >= oc.valid1) { left_strip = true; it++; break; }   // inside the image, outside this strip's valid rows
  }
  if (left_strip && it < rp.max_iter) {
    // undecided: the strip that owns row m continues at iteration `it` (sift_strip_resume)
    const int slot = atomicAdd(&ctr->n_left_strip, 1);
    if (walks && slot < walk_cap) {
      sift_walk w = w0;
      w.scaleLevel = s; w.y = m; w.x = nn; w.iteration = it;
      walks[slot] = w;
    }
    return;
  }
  // (a walk that jumped out on its last iteration never looks at the new sample: background.js:480 ends the loop)
  atomicAdd(&ctr->outcomes[outcome], 1);
  if (outcome == REFINE_ACCEPTED) {
    const int slot = atomicAdd(&ctr->n_kp, 1);
    if (slot < cap) out[slot] = kp;
  }
}

__global__ void __launch_bounds__(128)
refine_kernel(const OctaveDev *__restrict__ octs, const sift_candidate *__restrict__ cand,
              const int *__restrict__ d_ncand, int n_cand_host, int cand_cap, RefineParams rp,
              sift_keypoint *__restrict__ out, int cap, Counters *ctr, sift_walk *__restrict__ walks, int walk_cap)
{
  int n = (n_cand_host >= 0) ? n_cand_host : *d_ncand;
  if (n > cand_cap) n = cand_cap;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const sift_candidate cd = cand[idx];
    sift_walk w;
    w.octave = cd.octave; w.scaleLevel = cd.scaleLevel; w.x = cd.x; w.y = cd.y; w.iteration = 0; w.value = cd.value;
    w.candScale = cd.scaleLevel; w.candX = cd.x; w.candY = cd.y; w.reserved0 = 0;
    refine_walk(octs, w, rp, out, cap, ctr, walks, walk_cap);
  }
}

__global__ void __launch_bounds__(128)
refine_resume_kernel(const OctaveDev *__restrict__ octs, const sift_walk *__restrict__ in, int n, RefineParams rp,
                     sift_keypoint *__restrict__ out, int cap, Counters *ctr, sift_walk *__restrict__ walks, int walk_cap)
{
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x)
    refine_walk(octs, in[idx], rp, out, cap, ctr, walks, walk_cap);
}

void launch_refine(cudaStream_t st, const OctaveDev *d_octs, int n_octs, const sift_candidate *cand,
                   const int *d_ncand, int n_cand_host, int cand_cap, RefineParams rp, sift_keypoint *out, int cap,
                   Counters *ctr, sift_walk *walks, int walk_cap)
{
  (void)n_octs;
  int blocks = 148 * 4;
  if (n_cand_host >= 0) blocks = max(1, min(blocks, (n_cand_host + 127) / 128));
  refine_kernel<<<blocks, 128, 0, st>>>(d_octs, cand, d_ncand, n_cand_host, cand_cap, rp, out, cap, ctr, walks, walk_cap);
}

void launch_refine_resume(cudaStream_t st, const OctaveDev *d_octs, const sift_walk *in, int n, RefineParams rp,
                          sift_keypoint *out, int cap, Counters *ctr, sift_walk *walks, int walk_cap)
{
  if (n <= 0) return;
  refine_resume_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_octs, in, n, rp, out, cap, ctr, walks, walk_cap);
}

// ---- step functions SIFT_generateGradientVector / SIFT_generateHessianMatrix on Matrix2D (fp64) ----
struct DogView3 {
  const double *pl[3];
  size_t pitch;
  __device__ __forceinline__ double at(int s, int m, int n) const { return pl[s][(size_t)m * pitch + n]; }
};

__global__ void grad_hess_f64_kernel(const double *dm, const double *dc, const double *dp, int cols, int m, int n,
                                     double *out12)
{
  DogView3 D; D.pl[0] = dm; D.pl[1] = dc; D.pl[2] = dp; D.pitch = cols;
  double g[3], h[3][3];
  grad_hess(D, 1, m, n, g, h);
  for (int i = 0; i < 3; i++) out12[i] = g[i];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out12[3 + i * 3 + j] = h[i][j];
}

void launch_grad_hess_f64(cudaStream_t st, const double *dm, const double *dc, const double *dp, int cols,
                          int m, int n, double *out12)
{
  grad_hess_f64_kernel<<<1, 1, 0, st>>>(dm, dc, dp, cols, m, n, out12);
}
