// engine.cu -- context, planning (sigma schedule, radii, weights), device memory
// and the C ABI of include/sift_b200.h.  Host-side restatement of the reference's
// pipeline driver (background.js:71-237 octave/scale loop, :258 DoG loop, :359
// candidate loop, :455 refinement loop); the arithmetic lives in the kernels.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

#define SIFT_B200_VERSION "sift_b200 0.1.0 (sm_100a)"

static std::string g_create_error;

struct Scratch {
  void *p = nullptr;
  size_t cap = 0;
};

// One lane = one stream plus everything a frame in flight owns: its pyramid, candidate list, keypoint
// buffer (device + pinned mirror) and input image.  Images are independent, so frames of a
// batch are dealt round-robin to lanes and their kernels overlap: the small, latency-bound launches of
// the high octaves of one frame run next to the octave-0 launch of another, and uploads / downloads of
// one lane overlap the kernels of the others.  The stage API and single-image calls use lane 0.
#define SIFT_MAX_LANES 8
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_done = nullptr;
  Scratch planes, seeds, tbuf, image, cand, outbuf, tmaps, tmaps_t, tmaps_o0, tmaps_p0, order, walks;
  bool p0_maps = false;        // tmaps_p0 holds the TMA descriptors of octave 0's banded intermediate (blur_oct0p.cu)
  bool oct0_maps = false;      // tmaps_o0 holds the TMA-store descriptors of octave 0's planes (blur_oct0.cu)
  bool tma_scan = false;       // tmaps holds valid TMA descriptors of the DoG planes
  int tma_blur[SIFT_MAX_OCTAVES];   // per octave: maps of its T^T planes start at tmaps_t[tma_blur[o]] (-1: none)
  int cand_cap = 0, kp_cap = 0;
  void *h_out = nullptr;       // pinned mirror of outbuf
  size_t h_out_cap = 0;
  void *h_out2 = nullptr;      // second mirror: sift_detect_batch alternates, so a lane is re-issued while the host still
  size_t h_out2_cap = 0;       // orders its previous frame
  OctaveDev octs[SIFT_MAX_OCTAVES];
  OctaveDev *d_octs = nullptr;
  uint64_t plan_id = 0;        // plan the buffers above are laid out for
  bool busy = false;           // device work issued since the last join with the public stream
  // One frame of sift_detect_batch as a CUDA graph (counter reset, every blur launch, scan, refinement): the lane's
  // buffers do not move between frames, so the ~13 launches of a frame are replayed with one call.
  cudaGraphExec_t frame_graph = nullptr;
  uint64_t graph_plan = 0, graph_bufs = 0;      // plan / buffer generation the graph was captured for
  int graph_dtype = -1, graph_keep = -1, graph_launches = 0;
  sift_params graph_prm;
  uint64_t buf_version = 1;    // bumped whenever one of the lane's device buffers is (re)allocated
};

struct sift_ctx {
  int device = 0;
  cudaStream_t main_stream = nullptr;   // the public stream (sift_stream): lanes fork from / join into it
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  uint64_t bytes_h2d = 0, bytes_d2h = 0;   // image uploads / record downloads of the detect calls (sift_transfer_bytes)

  Lane lanes[SIFT_MAX_LANES];
  Lane *L = nullptr;                    // lane the stage helpers address
  double kp_running = 6500;             // decaying maximum of the keypoints per frame seen by sift_detect_batch (sizes the first download)
  bool use_graphs = true;               // SIFT_B200_NO_GRAPH=1: launch every kernel of a batch frame individually
  int n_lanes = 0, next_lane = 0;       // n_lanes == 0: chosen from the frame size (auto_lanes)

  // ---- optional per-kernel-class CUDA-event profiling (sift_set_profiling)
  bool profiling = false;
  struct ProfSpan { int kind; cudaEvent_t a, b; };
  std::vector<ProfSpan> spans;
  size_t spans_used = 0;

  // ---- plan (depends on params + input size)
  bool plan_valid = false;
  uint64_t plan_id = 0;
  sift_params prm;
  int in_w = 0, in_h = 0;
  int n_oct = 0, nlev = 0;
  int ow[SIFT_MAX_OCTAVES], oh[SIFT_MAX_OCTAVES];
  bool is_strip = false;                // the octave images are row strips of a mosaic (sift_strip_*)
  std::vector<sift_walk> escaped;       // refinement walks that left the strip in the last finish / resume
  sift_strip_layout strip;
  int strip_next_octave = 0;            // octave sift_strip_octave expects next
  int strip_dtype = 0;
  size_t strip_pitch = 0;
  // a strip's source rows are uploaded in STRIP_CHUNKS pieces on a copy stream; octave 0 runs band by band behind them
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t strip_ev[8] = {};
  int strip_chunks = 0, strip_chunk_rows = 0;             // 0 chunks: the upload went through the lane stream
  LevelPlan plans[SIFT_MAX_OCTAVES][SIFT_MAX_LEVELS];
  double dog_blur[SIFT_MAX_OCTAVES][SIFT_MAX_LEVELS];
  std::vector<double> h_weights;   // [0,256): u8 -> v/255.0 table, then per-level taps
  int poly_woff = 0;               // octave 0: merged polyphase tap table (blur_fused.cu)
  bool fused0 = false;             // octave 0 runs the fused polyphase kernel
  int mma0_woff = -1;              // octave 0: per-lane band fragments of the DMMA kernel (blur_mma.cu), -1: not usable
  bool no_mma = false;             // SIFT_B200_NO_MMA=1: scalar-FMA blur kernels everywhere
  bool mma_big_only = false;       // SIFT_B200_MMA_BIG_ONLY=1: octaves below 2^18 pixels keep the scalar two-pass kernels
  bool oct0_variant_forced = false;  // one of the octave-0 variant knobs is set: it wins over the DMMA kernel
  bool sep_variant_forced = false;   // one of the knobs of the scalar two-pass kernels is set: they win over the DMMA passes
  bool force_generic = false;      // SIFT_B200_FORCE_GENERIC=1: radius-generic two-pass kernels everywhere
  bool no_tma = false;             // SIFT_B200_NO_TMA=1: the pointer-chasing scan instead of the TMA-tiled one
  bool force_old = false;          // SIFT_B200_FORCE_OLD=1: the row-major-T fallback kernels (blur_generic.cu) everywhere
  double *d_weights = nullptr;
  size_t d_weights_cap = 0;

  // ---- lane-0-only arenas of the stage / step API (grow only)
  Scratch low, misc[6];
  int low_cap = 0;
  std::vector<uint64_t> sort_a, sort_b;
  void *h_cand = nullptr;     // pinned candidate staging
  size_t h_cand_cap = 0;

  bool pyramid_built = false;
  uint64_t pyramid_serial = 0;          // bumped whenever the pyramid the stage API reads is rebuilt, replaced or invalidated
  int keep_gauss = 1;
  Counters last;
};

// ------------------------------------------------------------------ helpers ----
static int fail(sift_ctx *c, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, SIFT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static int grow(sift_ctx *ctx, Scratch &s, size_t bytes)
{
  if (bytes <= s.cap) return SIFT_OK;
  if (s.p) { CK(cudaStreamSynchronize(ctx->L->stream)); CK(cudaFree(s.p)); s.p = nullptr; s.cap = 0; }
  bytes = (bytes + 255) & ~(size_t)255;
  CK(cudaMalloc(&s.p, bytes));
  s.cap = bytes;
  ctx->L->buf_version++;                 // captured frame graphs of this lane point at the old buffers
  return SIFT_OK;
}

static int grow_pinned(sift_ctx *ctx, void **p, size_t *cap, size_t bytes)
{
  if (bytes <= *cap) return SIFT_OK;
  if (*p) { CK(cudaStreamSynchronize(ctx->L->stream)); CK(cudaFreeHost(*p)); *p = nullptr; *cap = 0; }
  CK(cudaMallocHost(p, bytes));
  *cap = bytes;
  return SIFT_OK;
}

// Per-kernel-class timing: a pair of events around each launch group, summed at read-out.
static void prof_begin(sift_ctx *ctx, int kind)
{
  if (!ctx->profiling) return;
  if (ctx->spans_used == ctx->spans.size()) {
    sift_ctx::ProfSpan sp; sp.kind = kind;
    cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
    ctx->spans.push_back(sp);
  }
  ctx->spans[ctx->spans_used].kind = kind;
  cudaEventRecord(ctx->spans[ctx->spans_used].a, ctx->L->stream);
}
static void prof_end(sift_ctx *ctx)
{
  if (!ctx->profiling) return;
  cudaEventRecord(ctx->spans[ctx->spans_used].b, ctx->L->stream);
  ctx->spans_used++;
}

static double js_round(double x) { double f = std::floor(x); return (x - f >= 0.5) ? f + 1.0 : f; }

static size_t dtype_size(int dtype)
{
  switch (dtype) {
    case SIFT_U8: return 1;
    case SIFT_F32: return 4;
    case SIFT_F64: return 8;
    case SIFT_RGBA8: return 4;
  }
  return 0;
}

// Normalised 1D Gaussian taps w[0..2R]: the reference's 2D kernel (sift.js:31-67) is
// g(i,j)/sum with g = exp(-(i^2+j^2)/(2 sigma^2))/(2 pi sigma^2) = outer product of these.
static void gaussian_taps(double sigma, int R, double *w)
{
  double s = 0;
  for (int i = 0; i <= 2 * R; i++) {
    const double d = (double)(i - R);
    w[i] = std::exp(((d * d) / (sigma * sigma)) * -0.5);
    s += w[i];
  }
  for (int i = 0; i <= 2 * R; i++) w[i] /= s;
}

static bool same_params(const sift_params &a, const sift_params &b)
{
  return a.numberOfOctaves == b.numberOfOctaves && a.scalesPerOctave == b.scalesPerOctave &&
         a.minBlurLevel == b.minBlurLevel && a.assumedBlur == b.assumedBlur;
}

// background.js:89-177: blur levels, offset sigmas and radii of every level (host only, no device needed).
// Returns 0 or the (octave * 100 + level) of the level whose offset sigma is not realisable.
static int level_schedule(const sift_params *p, LevelPlan plans[SIFT_MAX_OCTAVES][SIFT_MAX_LEVELS])
{
  const int spo = p->scalesPerOctave, nlev = spo + 3, n_oct = p->numberOfOctaves;
  const double k = std::pow(2.0, 1.0 / spo);                                   // background.js:100
  double base_blur = p->minBlurLevel;                                          // background.js:89
  for (int o = 0; o < n_oct; o++)
    for (int s = 0; s < nlev; s++) {
      LevelPlan &lp = plans[o][s];
      lp.woff = -1;
      if (o > 0 && s == 0) {
        base_blur = plans[o - 1][spo].blurLevel;                               // background.js:122
        lp.blurLevel = base_blur; lp.offsetSigma = 0; lp.radius = 0;
        continue;
      }
      const double current_k = std::pow(k, (double)s);                         // :157
      const double target = base_blur * current_k;                             // :173
      const double base_sigma = (o == 0) ? p->assumedBlur : base_blur;         // :174-176
      const double off = std::sqrt((target * target) - (base_sigma * base_sigma));   // :177
      lp.blurLevel = target; lp.offsetSigma = off;
      if (!(off > 0) || !std::isfinite(off)) return o * 100 + s + 1;
      lp.radius = (int)js_round(3 * off);                                      // sift.js:38
    }
  return 0;
}

static int check_params(sift_ctx *ctx, const sift_params *p)
{
  if (!p) return fail(ctx, SIFT_ERR_BAD_ARGS, "params is NULL");
  if (p->numberOfOctaves < 1 || p->numberOfOctaves > SIFT_MAX_OCTAVES)
    return fail(ctx, SIFT_ERR_UNSUPPORTED, "numberOfOctaves %d outside 1..%d", p->numberOfOctaves, SIFT_MAX_OCTAVES);
  if (p->scalesPerOctave < 1 || p->scalesPerOctave + 3 > SIFT_MAX_LEVELS)
    return fail(ctx, SIFT_ERR_UNSUPPORTED, "scalesPerOctave %d outside 1..%d", p->scalesPerOctave, SIFT_MAX_LEVELS - 3);
  return SIFT_OK;
}

// Row strips of a mosaic.  All rows are rows of the GLOBAL octave grids.
static int strip_layout(sift_ctx *ctx, const sift_params *p, int full_w, int full_h, int row0, int row1, int margin,
                        sift_strip_layout *out)
{
  int rc;
  if ((rc = check_params(ctx, p))) return rc;
  if (!out || full_w < 1 || full_h < 1 || margin < 1) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad mosaic size / margin");
  const int n_oct = p->numberOfOctaves, nlev = p->scalesPerOctave + 3;
  const int H0 = 2 * full_h, align = 1 << (n_oct - 1);
  if (row0 < 0 || row1 <= row0 || row1 > H0 || (row0 % align) != 0 || (row1 != H0 && (row1 % align) != 0))
    return fail(ctx, SIFT_ERR_BAD_ARGS, "strip rows [%d, %d) must lie in [0, %d) on multiples of %d (2^(octaves-1))",
                row0, row1, H0, align);
  LevelPlan plans[SIFT_MAX_OCTAVES][SIFT_MAX_LEVELS];
  const int bad = level_schedule(p, plans);
  if (bad) return fail(ctx, SIFT_ERR_BAD_ARGS, "sigma schedule not realisable at octave %d level %d", (bad - 1) / 100, (bad - 1) % 100);
  memset(out, 0, sizeof *out);
  out->octaves = n_oct;
  int W = 2 * full_w, H = H0;
  for (int o = 0; o < n_oct; o++) {
    int rmax = 0;
    for (int s = 0; s < nlev; s++) rmax = std::max(rmax, plans[o][s].radius);
    const int halo = (rmax + margin + 1) & ~1;
    const int own0 = row0 >> o, own1 = (row1 == H0) ? H : (row1 >> o);
    out->width[o] = W; out->height[o] = H;
    out->own0[o] = own0; out->own1[o] = own1;
    out->top[o] = std::max(0, own0 - halo) & ~1;
    out->bottom[o] = std::min(H, own1 + halo);
    out->halo[o] = halo;
    if (own1 <= own0) return fail(ctx, SIFT_ERR_UNSUPPORTED, "strip owns no row of octave %d", o);
    // the halo comes from the direct neighbours only: a strip must be at least one halo tall in every octave
    if ((row0 > 0 || row1 < H0) && own1 - own0 < halo)
      return fail(ctx, SIFT_ERR_UNSUPPORTED, "strip is %d rows tall in octave %d but the halo is %d: gather this octave instead",
                  own1 - own0, o, halo);
    W = (W + 1) / 2; H = (H + 1) / 2;
  }
  return SIFT_OK;
}

// background.js:84-177: sizes, blur levels, offset sigmas, radii; allocates the pyramid.
// background.js:84-177: sizes, blur levels, offset sigmas, radii; allocates the pyramid.
static int ensure_plan(sift_ctx *ctx, int w, int h, const sift_params *p, const sift_strip_layout *strip = nullptr)
{
  if (!p) return fail(ctx, SIFT_ERR_BAD_ARGS, "params is NULL");
  if (w < 1 || h < 1) return fail(ctx, SIFT_ERR_BAD_ARGS, "image size %dx%d", w, h);
  if (p->numberOfOctaves < 1 || p->numberOfOctaves > SIFT_MAX_OCTAVES)
    return fail(ctx, SIFT_ERR_UNSUPPORTED, "numberOfOctaves %d outside 1..%d", p->numberOfOctaves, SIFT_MAX_OCTAVES);
  if (p->scalesPerOctave < 1 || p->scalesPerOctave + 3 > SIFT_MAX_LEVELS)
    return fail(ctx, SIFT_ERR_UNSUPPORTED, "scalesPerOctave %d outside 1..%d", p->scalesPerOctave, SIFT_MAX_LEVELS - 3);
  if ((int64_t)w * 2 > (1 << 30) || (int64_t)h * 2 > (1 << 30))
    return fail(ctx, SIFT_ERR_UNSUPPORTED, "image too large");
  const bool same_strip = (!strip && !ctx->is_strip) || (strip && ctx->is_strip && memcmp(strip, &ctx->strip, sizeof *strip) == 0);
  if (ctx->plan_valid && ctx->in_w == w && ctx->in_h == h && same_params(ctx->prm, *p) && same_strip) {
    ctx->prm = *p;   // thresholds may differ; they do not affect the plan
    return SIFT_OK;
  }
  ctx->plan_valid = false;
  ctx->pyramid_built = false;
  ctx->pyramid_serial++;
  const int spo = p->scalesPerOctave, nlev = spo + 3, n_oct = p->numberOfOctaves;
  (void)spo;

  // sizes: octave 0 is the 2x nearest-neighbour upsample (background.js:84), then ceil halves (matrix2d.js:119)
  int ow[SIFT_MAX_OCTAVES], oh[SIFT_MAX_OCTAVES];
  ow[0] = 2 * w; oh[0] = 2 * h;
  for (int o = 1; o < n_oct; o++) { ow[o] = (ow[o - 1] + 1) / 2; oh[o] = (oh[o - 1] + 1) / 2; }

  if (strip) {                          // octave images are the strip's rows [top, bottom) of the global octaves
    for (int o = 0; o < n_oct; o++) { ow[o] = strip->width[o]; oh[o] = strip->bottom[o] - strip->top[o]; }
  }
  const int bad = level_schedule(p, ctx->plans);
  if (bad)
    return fail(ctx, SIFT_ERR_BAD_ARGS, "sigma schedule not realisable at octave %d level %d (target %g)", (bad - 1) / 100,
                (bad - 1) % 100, ctx->plans[(bad - 1) / 100][(bad - 1) % 100].blurLevel);
  ctx->h_weights.clear();
  for (int v = 0; v < 256; v++) ctx->h_weights.push_back((double)v / 255.0);   // image-utils.js:114
  for (int o = 0; o < n_oct; o++) {
    for (int s = 0; s < nlev; s++) {
      LevelPlan &lp = ctx->plans[o][s];
      if (o > 0 && s == 0) continue;
      if (lp.radius > 4096) return fail(ctx, SIFT_ERR_UNSUPPORTED, "kernel radius %d too large", lp.radius);
      lp.woff = (int)ctx->h_weights.size();
      ctx->h_weights.resize(ctx->h_weights.size() + 2 * lp.radius + 1 + SIFT_WPAD, 0.0);
      gaussian_taps(lp.offsetSigma, lp.radius, ctx->h_weights.data() + lp.woff);
    }
    for (int s = 1; s < nlev; s++) ctx->dog_blur[o][s - 1] = ctx->plans[o][s - 1].blurLevel;   // background.js:327
  }
  ctx->fused0 = !ctx->force_generic && fused0_supported(ctx->plans[0], nlev);
  if (ctx->fused0) {
    const int per = fused0_taps_per_level();
    ctx->poly_woff = (int)ctx->h_weights.size();
    ctx->h_weights.resize(ctx->h_weights.size() + (size_t)nlev * per, 0.0);
    for (int s = 0; s < nlev; s++) {
      const LevelPlan &lp = ctx->plans[0][s];
      fused0_merge_taps(ctx->h_weights.data() + lp.woff, lp.radius, ctx->h_weights.data() + ctx->poly_woff + (size_t)s * per);
    }
  }
  ctx->mma0_woff = -1;
  if (ctx->fused0 && !ctx->no_mma && mma0_supported(ctx->plans[0], nlev)) {
    const int per = fused0_taps_per_level(), fper = mma0_frag_doubles(1);
    ctx->mma0_woff = (int)ctx->h_weights.size();
    ctx->h_weights.resize(ctx->h_weights.size() + (size_t)mma0_frag_doubles(nlev), 0.0);
    for (int s = 0; s < nlev; s++)
      mma0_build_frags(ctx->h_weights.data() + ctx->poly_woff + (size_t)s * per, ctx->plans[0][s].radius,
                       ctx->h_weights.data() + ctx->mma0_woff + (size_t)s * fper);
  }

  for (int o = 0; o < n_oct; o++) { ctx->ow[o] = ow[o]; ctx->oh[o] = oh[o]; }
  // ---- weights (shared by all lanes)
  const size_t wbytes = ctx->h_weights.size() * sizeof(double);
  CK(cudaDeviceSynchronize());          // a plan change is rare; no lane may still read the old tables
  if (wbytes > ctx->d_weights_cap) {
    if (ctx->d_weights) CK(cudaFree(ctx->d_weights));
    CK(cudaMalloc((void **)&ctx->d_weights, wbytes));
    ctx->d_weights_cap = wbytes;
  }
  CK(cudaMemcpy(ctx->d_weights, ctx->h_weights.data(), wbytes, cudaMemcpyHostToDevice));
  ctx->prm = *p; ctx->in_w = w; ctx->in_h = h; ctx->n_oct = n_oct; ctx->nlev = nlev;
  ctx->is_strip = strip != nullptr;
  if (strip) ctx->strip = *strip;
  ctx->plan_id++;
  ctx->plan_valid = true;
  return SIFT_OK;
}

// Device memory of one lane for the current plan (lazy: lanes other than 0 are only touched by batches).
static int ensure_lane(sift_ctx *ctx, Lane *ln)
{
  if (ln->plan_id == ctx->plan_id) return SIFT_OK;
  Lane *saved = ctx->L;
  ctx->L = ln;                          // grow() synchronises ctx->L->stream before freeing
  const int n_oct = ctx->n_oct, nlev = ctx->nlev, w = ctx->in_w, h = ctx->in_h;
  const int *ow = ctx->ow, *oh = ctx->oh;
  size_t plane_elems = 0, seed_elems = 0, t_bytes = 0;
  for (int o = 0; o < n_oct; o++) {
    const int pitch = (ow[o] + 31) & ~31;
    plane_elems += (size_t)(2 * nlev - 1) * oh[o] * pitch;
    if (o > 0) seed_elems += (size_t)ow[o] * oh[o];
    const size_t trows = (o == 0) ? (size_t)h : (size_t)oh[o];
    const size_t tl = (o == 0) ? nlev : nlev - 1;
    if (o > 0 || !ctx->fused0)
      t_bytes = std::max(t_bytes, std::max(tl * trows * ow[o], sep_t_elems(ow[o], oh[o], (int)trows, ctx->plans[o], o == 0 ? 0 : 1, nlev)) * sizeof(double));
    if (o > 0 && !ctx->no_mma && mma_sep_supported(ctx->plans[o], nlev, ow[o], oh[o]))
      t_bytes = std::max(t_bytes, mma_sep_t_elems(ctx->plans[o], nlev, ow[o], oh[o]) * sizeof(double));
  }
  const bool want_p0 = ctx->fused0 && !ctx->no_tma && oct0p_supported(ctx->plans[0], nlev);
  if (want_p0) t_bytes = std::max(t_bytes, oct0p_t_bytes(w, h, nlev));
  int rc = SIFT_OK;
  do {
    if ((rc = grow(ctx, ln->planes, plane_elems * sizeof(float)))) break;
    if ((rc = grow(ctx, ln->seeds, std::max<size_t>(seed_elems, 1) * sizeof(double)))) break;
    if ((rc = grow(ctx, ln->tbuf, std::max<size_t>(t_bytes, 8)))) break;
    float *pp = (float *)ln->planes.p;
    double *sp = (double *)ln->seeds.p;
    for (int o = 0; o < n_oct; o++) {
      OctaveDev &od = ln->octs[o];
      memset(&od, 0, sizeof od);
      od.w = ow[o]; od.h = oh[o]; od.pitch = (ow[o] + 31) & ~31; od.nlev = nlev;
      od.y_top = 0; od.gh = oh[o]; od.own0 = 0; od.own1 = oh[o]; od.seed_off = 0;
      od.valid0 = 0; od.valid1 = oh[o];
      if (ctx->is_strip) {
        const sift_strip_layout &sl = ctx->strip;
        od.y_top = sl.top[o]; od.gh = sl.height[o]; od.own0 = sl.own0[o] - sl.top[o]; od.own1 = sl.own1[o] - sl.top[o];
        od.seed_off = (o + 1 < n_oct) ? sl.top[o] / 2 - sl.top[o + 1] : 0;
        int rmax = 0;
        for (int s = 0; s < nlev; s++) rmax = std::max(rmax, ctx->plans[o][s].radius);
        od.valid0 = sl.top[o] == 0 ? 0 : std::min(rmax, od.own0);
        od.valid1 = sl.bottom[o] == sl.height[o] ? oh[o] : std::max(oh[o] - rmax, od.own1);
      }
      const size_t pe = (size_t)od.h * od.pitch;
      for (int s = 0; s < nlev; s++) { od.gauss[s] = pp; pp += pe; }
      for (int s = 0; s < nlev - 1; s++) { od.dog[s] = pp; pp += pe; }
      if (o > 0) { od.seed64 = sp; sp += (size_t)od.w * od.h; }
    }
    if (!ln->d_octs && cudaMalloc((void **)&ln->d_octs, sizeof(OctaveDev) * SIFT_MAX_OCTAVES) != cudaSuccess) {
      rc = fail(ctx, SIFT_ERR_CUDA, "cudaMalloc(d_octs) failed");
      break;
    }
    // pageable source: staged before the call returns, so octs may change afterwards
    if (cudaMemcpyAsync(ln->d_octs, ln->octs, sizeof(OctaveDev) * n_oct, cudaMemcpyHostToDevice, ln->stream) != cudaSuccess) {
      rc = fail(ctx, SIFT_ERR_CUDA, "upload of the octave table failed");
      break;
    }
    // TMA descriptors of the DoG planes for the tiled scan
    ln->tma_scan = false;
    if (!ctx->no_tma && scan_tma_supported(nlev - 1)) {
      std::vector<char> hm(scan_tma_map_bytes(n_oct, nlev - 1));
      if (scan_tma_build_maps(ln->octs, n_oct, nlev - 1, hm.data())) {
        if ((rc = grow(ctx, ln->tmaps, hm.size()))) break;
        if (cudaMemcpyAsync(ln->tmaps.p, hm.data(), hm.size(), cudaMemcpyHostToDevice, ln->stream) != cudaSuccess ||
            cudaStreamSynchronize(ln->stream) != cudaSuccess) {
          rc = fail(ctx, SIFT_ERR_CUDA, "upload of the TMA descriptors failed");
          break;
        }
        ln->tma_scan = true;
      }
    }
    // TMA descriptors of octave 0's banded intermediate (blur_oct0p.cu); the planes alias tbuf
    ln->p0_maps = false;
    if (want_p0) {
      std::vector<char> hm(oct0p_map_bytes(nlev));
      if (oct0p_build_maps((double *)ln->tbuf.p, w, h, nlev, hm.data())) {
        if ((rc = grow(ctx, ln->tmaps_p0, hm.size()))) break;
        if (cudaMemcpyAsync(ln->tmaps_p0.p, hm.data(), hm.size(), cudaMemcpyHostToDevice, ln->stream) != cudaSuccess ||
            cudaStreamSynchronize(ln->stream) != cudaSuccess) {
          rc = fail(ctx, SIFT_ERR_CUDA, "upload of the octave-0 band TMA descriptors failed");
          break;
        }
        ln->p0_maps = true;
      }
    }
    // TMA-store descriptors of octave 0's Gaussian / DoG planes (blur_oct0.cu)
    ln->oct0_maps = false;
    if (ctx->fused0 && oct0_v2_supported(ctx->plans[0], nlev)) {
      std::vector<char> hm(oct0_out_map_bytes(nlev));
      if (oct0_build_out_maps(ln->octs[0], nlev, hm.data())) {
        if ((rc = grow(ctx, ln->tmaps_o0, hm.size()))) break;
        if (cudaMemcpyAsync(ln->tmaps_o0.p, hm.data(), hm.size(), cudaMemcpyHostToDevice, ln->stream) != cudaSuccess ||
            cudaStreamSynchronize(ln->stream) != cudaSuccess) {
          rc = fail(ctx, SIFT_ERR_CUDA, "upload of the octave-0 TMA descriptors failed");
          break;
        }
        ln->oct0_maps = true;
      }
    }
    // TMA descriptors of the fp64 T^T planes (pass B of octaves >= 1); the planes of every octave alias tbuf
    for (int o = 0; o < SIFT_MAX_OCTAVES; o++) ln->tma_blur[o] = -1;
    if (!ctx->no_tma) {
      std::vector<char> hm(sep_tma_map_bytes(n_oct * nlev));
      int used = 0;
      for (int o = 1; o < n_oct; o++) {
        if (!sep_supported(ctx->plans[o], 1, nlev, ow[o], oh[o])) continue;
        const int n = sep_tma_build_maps(ctx->plans[o], 1, nlev, (double *)ln->tbuf.p, ow[o], oh[o], oh[o],
                                         hm.data() + sep_tma_map_bytes(used));
        if (n > 0) { ln->tma_blur[o] = used; used += n; }
      }
      if (used > 0) {
        if ((rc = grow(ctx, ln->tmaps_t, hm.size()))) break;
        if (cudaMemcpyAsync(ln->tmaps_t.p, hm.data(), sep_tma_map_bytes(used), cudaMemcpyHostToDevice, ln->stream) != cudaSuccess ||
            cudaStreamSynchronize(ln->stream) != cudaSuccess) {
          rc = fail(ctx, SIFT_ERR_CUDA, "upload of the blur TMA descriptors failed");
          break;
        }
      }
    }
    // candidate / keypoint capacity: extrema are ~3e-4 of the voxels on the synthetic frames
    const int want = (int)std::min<int64_t>(std::max<int64_t>(16384, ((int64_t)w * h) / 8), 1 << 26);
    if (want > ln->cand_cap) {
      if ((rc = grow(ctx, ln->cand, (size_t)want * sizeof(sift_candidate)))) break;
      ln->cand_cap = want;
    }
    if (want > ln->kp_cap) {
      if ((rc = grow(ctx, ln->outbuf, sizeof(Counters) + (size_t)want * sizeof(sift_keypoint)))) break;
      ln->kp_cap = want;
    }
    ln->plan_id = ctx->plan_id;
  } while (0);
  ctx->L = saved;
  return rc;
}

// plan + lane 0 (what every single-image / stage entry point needs)
static int ensure_plan0(sift_ctx *ctx, int w, int h, const sift_params *p)
{
  int rc;
  ctx->L = &ctx->lanes[0];
  if ((rc = ensure_plan(ctx, w, h, p))) return rc;
  return ensure_lane(ctx, ctx->L);
}

static Counters *dev_counters(sift_ctx *ctx) { return (Counters *)ctx->L->outbuf.p; }
static sift_keypoint *dev_keypoints(sift_ctx *ctx) { return (sift_keypoint *)((char *)ctx->L->outbuf.p + sizeof(Counters)); }

// Blur + DoG of one octave and the seed of the next, from an image already on the device (octave 0) or from
// the octave's seed (octaves >= 1).
static void run_octave(sift_ctx *ctx, int o, const void *d_image, int dtype, size_t pitch_bytes)
{
  const int spo = ctx->prm.scalesPerOctave;
  cudaStream_t st = ctx->L->stream;
  const OctaveDev &od = ctx->L->octs[o];
  const OctaveDev *next = (o + 1 < ctx->n_oct) ? &ctx->L->octs[o + 1] : nullptr;
  if (o == 0 && ctx->fused0) {
    prof_begin(ctx, SIFT_PROF_BLUR_OCT0);
    if (ctx->mma0_woff >= 0 && !ctx->oct0_variant_forced &&
        launch_oct0_mma(st, d_image, dtype, pitch_bytes, ctx->in_w, ctx->in_h, od, next, ctx->d_weights + ctx->mma0_woff,
                        ctx->plans[0], ctx->nlev, spo, ctx->keep_gauss)) {
      // done
    } else if (ctx->L->p0_maps && !ctx->L->oct0_maps && !oct0_small_supported(ctx->plans[0], ctx->nlev)) {
      ctx->launches += launch_oct0p(st, d_image, dtype, pitch_bytes, ctx->in_w, ctx->in_h, od, next, ctx->d_weights, ctx->plans[0],
                                    ctx->poly_woff, ctx->nlev, spo, ctx->keep_gauss, (double *)ctx->L->tbuf.p, ctx->L->tmaps_p0.p) - 1;
    } else if (!ctx->L->oct0_maps && oct0_small_supported(ctx->plans[0], ctx->nlev))
      launch_oct0_small(st, d_image, dtype, pitch_bytes, ctx->in_w, ctx->in_h, od, next, ctx->d_weights, ctx->plans[0],
                        ctx->poly_woff, ctx->nlev, spo, ctx->keep_gauss);
    else if (ctx->L->oct0_maps)
      launch_oct0_v2(st, d_image, dtype, pitch_bytes, ctx->in_w, ctx->in_h, od, next, ctx->d_weights, ctx->plans[0],
                     ctx->poly_woff, ctx->nlev, spo, ctx->keep_gauss, ctx->L->tmaps_o0.p);
    else
      launch_fused_octave0(st, d_image, dtype, pitch_bytes, ctx->in_w, ctx->in_h, od, next, ctx->d_weights, ctx->plans[0],
                           ctx->poly_woff, ctx->nlev, spo, ctx->keep_gauss, ctx->d_weights);
    ctx->launches += 1;
    prof_end(ctx);
    return;
  }
  const int first = (o == 0) ? 0 : 1;
  const int hrows = (o == 0) ? ctx->in_h : od.h;
  prof_begin(ctx, o == 0 ? SIFT_PROF_BLUR_OCT0 : (o == 1 ? SIFT_PROF_BLUR_OCT1 : SIFT_PROF_BLUR_HIGH));
  // small octaves (< 2^18 pixels: 36-75 CTAs) run 20 % longer on the DMMA passes than on the scalar ones when a frame is
  // alone on the GPU, but with frames in flight the whole path is 2-5 % faster with them (they leave more of the SMs to
  // the other frames' kernels): DMMA everywhere; SIFT_B200_MMA_BIG_ONLY=1 keeps the scalar passes for small octaves
  if (o > 0 && !ctx->no_mma && !ctx->force_old && !ctx->force_generic && !ctx->sep_variant_forced &&
      ((long long)od.w * od.h >= (1 << 18) || !ctx->mma_big_only) && mma_sep_supported(ctx->plans[o], ctx->nlev, od.w, od.h)) {
    launch_mma_sep(st, od, next, ctx->d_weights, ctx->plans[o], (double *)ctx->L->tbuf.p, spo, ctx->keep_gauss);
    ctx->launches += 2;
    prof_end(ctx);
    return;
  }
  if (!ctx->force_old && sep_supported(ctx->plans[o], first, ctx->nlev, od.w, od.h)) {
    double *tb = (double *)ctx->L->tbuf.p;
    const void *tm = (o > 0 && ctx->L->tma_blur[o] >= 0) ? (const char *)ctx->L->tmaps_t.p + sep_tma_map_bytes(ctx->L->tma_blur[o]) : nullptr;
    if (o == 0)
      launch_sep_pass_a(st, d_image, dtype, pitch_bytes, ctx->in_w, 1, od.w, hrows, od.h, ctx->d_weights,
                        ctx->plans[o], first, ctx->nlev, tb);
    else
      launch_sep_pass_a(st, od.seed64, SIFT_F64, (size_t)od.w * sizeof(double), od.w, 0, od.w, hrows, od.h,
                        ctx->d_weights, ctx->plans[o], first, ctx->nlev, tb);
    launch_sep_pass_b(st, o == 0, od, ctx->d_weights, ctx->plans[o], first, tb, hrows, next, spo, ctx->keep_gauss, tm);
  } else {
    // radii too large for the staged tiles (octaves >= 4): row-major T, level-parallel kernels
    double *T[SIFT_MAX_LEVELS];
    for (int i = 0; i < ctx->nlev - first; i++) T[i] = (double *)ctx->L->tbuf.p + (size_t)i * hrows * od.w;
    if (o == 0)
      launch_hblur(st, d_image, dtype, pitch_bytes, ctx->in_w, ctx->in_h, 1, od.w, hrows, ctx->d_weights,
                   ctx->plans[o], first, ctx->nlev, T, nullptr);
    else
      launch_hblur(st, od.seed64, SIFT_F64, (size_t)od.w * sizeof(double), od.w, od.h, 0, od.w, hrows,
                   ctx->d_weights, ctx->plans[o], first, ctx->nlev, T, nullptr);
    launch_vblur(st, o == 0, od, ctx->d_weights, ctx->plans[o], first, T, next, spo, ctx->keep_gauss);
  }
  ctx->launches += 2;
  prof_end(ctx);
}

// Gaussian scale space + DoG + seeds for every octave, from an image already on the device.
static int run_pyramid(sift_ctx *ctx, const void *d_image, int dtype, size_t pitch_bytes)
{
  ctx->pyramid_serial++;
  for (int o = 0; o < ctx->n_oct; o++) run_octave(ctx, o, d_image, dtype, pitch_bytes);
  CK(cudaGetLastError());
  ctx->pyramid_built = true;
  return SIFT_OK;
}

static double contrast_threshold(const sift_params &p)
{
  // sift.js:285 / background.js:572
  return ((std::pow(2.0, 1.0 / p.scalesPerOctave) - 1) / (std::pow(2.0, 1.0 / 3) - 1)) * p.contrastThreshold;
}

static int run_scan(sift_ctx *ctx, int count_low, const sift_params *thr = nullptr)
{
  sift_params tp = ctx->prm;
  if (thr) { tp.contrastThreshold = thr->contrastThreshold; tp.preFilterFactor = thr->preFilterFactor; }
  const double pix_thr = contrast_threshold(tp) * tp.preFilterFactor;               // sift.js:293
  prof_begin(ctx, SIFT_PROF_SCAN);
  if (ctx->L->tma_scan)
    launch_scan_tma(ctx->L->stream, ctx->L->octs, ctx->L->tmaps.p, ctx->n_oct, ctx->prm.scalesPerOctave, pix_thr, count_low,
                    (sift_candidate *)ctx->L->cand.p, ctx->L->cand_cap, (sift_candidate *)ctx->low.p, ctx->low_cap,
                    dev_counters(ctx));
  else
    launch_scan_all(ctx->L->stream, ctx->L->octs, ctx->L->d_octs, ctx->n_oct, ctx->prm.scalesPerOctave, pix_thr, count_low,
                    (sift_candidate *)ctx->L->cand.p, ctx->L->cand_cap, (sift_candidate *)ctx->low.p, ctx->low_cap,
                    dev_counters(ctx));
  ctx->launches += 1;
  prof_end(ctx);
  CK(cudaGetLastError());
  return SIFT_OK;
}

static RefineParams refine_params(const sift_ctx *ctx, const sift_params *over)
{
  sift_params p = ctx->prm;
  if (over) {
    p.contrastThreshold = over->contrastThreshold; p.edgeRatio = over->edgeRatio;
    p.maxIterations = over->maxIterations; p.offsetBound = over->offsetBound;
    p.minBlurLevel = over->minBlurLevel; p.minInterpixelDistance = over->minInterpixelDistance;
  }
  RefineParams rp;
  rp.spo = ctx->prm.scalesPerOctave;
  rp.ndog = ctx->nlev - 1;
  rp.max_iter = p.maxIterations;
  rp.offset_bound = p.offsetBound;
  rp.contrast_thr = contrast_threshold(p);
  rp.edge_thr = ((p.edgeRatio + 1) * (p.edgeRatio + 1)) / p.edgeRatio;   // background.js:598
  rp.min_blur = p.minBlurLevel;
  rp.min_interpixel = p.minInterpixelDistance;
  return rp;
}

#define WALK_CAP 65536       // refinement walks leaving a strip per finish / resume call (a handful in practice)
static int run_refine(sift_ctx *ctx, int n_cand_host, sift_keypoint *d_out, int cap, const sift_params *over = nullptr)
{
  prof_begin(ctx, SIFT_PROF_REFINE);
  if (ctx->is_strip) {                                     // walks that leave the strip are recorded, not decided
    int rc = grow(ctx, ctx->L->walks, 2 * (size_t)WALK_CAP * sizeof(sift_walk));   // [0, CAP): out, [CAP, 2 CAP): resume input
    if (rc) return rc;
  }
  launch_refine(ctx->L->stream, ctx->L->d_octs, ctx->n_oct, (const sift_candidate *)ctx->L->cand.p,
                &dev_counters(ctx)->n_cand, n_cand_host, ctx->L->cand_cap, refine_params(ctx, over), d_out, cap,
                dev_counters(ctx), ctx->is_strip ? (sift_walk *)ctx->L->walks.p : nullptr, WALK_CAP);
  prof_end(ctx);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return SIFT_OK;
}

static void fill_stats(sift_stats *s, const Counters &c, int count_low, float ms, int launches)
{
  if (!s) return;
  s->candidates = c.n_cand;
  s->lowContrastExtrema = count_low ? c.n_low : -1;
  s->keypoints = c.outcomes[REFINE_ACCEPTED];
  s->rejLowContrast = c.outcomes[REFINE_LOW_CONTRAST];
  s->rejEdge = c.outcomes[REFINE_EDGE];
  s->rejLeftScale = c.outcomes[REFINE_LEFT_SCALE];
  s->rejLeftRows = c.outcomes[REFINE_LEFT_ROWS];
  s->rejLeftCols = c.outcomes[REFINE_LEFT_COLS];
  s->rejNoConvergence = c.outcomes[REFINE_NO_CONVERGENCE];
  s->rejSingular = c.outcomes[REFINE_SINGULAR];
  s->msDevice = ms;
  s->kernelLaunches = launches;
  s->leftStrip = c.n_left_strip;
}

static inline uint64_t cand_key(int o, int s, int y, int x)
{
  return ((uint64_t)(uint32_t)o << 58) | ((uint64_t)(uint32_t)s << 52) | ((uint64_t)(uint32_t)y << 26) | (uint64_t)(uint32_t)x;
}

// Reference output order = candidate order: octave, scale, row, column (background.js:468-471, sift.js:221-222).
// LSD radix sort of (compact key, index) and a gather into `dst` (may alias nothing in `src`).  The key packs
// only as many bits as the pyramid needs (octave-0 size), so a 1080p frame sorts in three 11-bit passes.
static void sort_keypoints_into(sift_ctx *ctx, const sift_keypoint *src, int n, sift_keypoint *dst, int dst_cap)
{
  if (n <= 0) return;
  int bx = 1, by = 1;
  while ((1 << bx) < ctx->ow[0]) bx++;
  const int gh0 = ctx->is_strip ? ctx->strip.height[0] : ctx->oh[0];      // candY is a row of the whole image
  while ((1 << by) < gh0) by++;
  std::vector<uint64_t> &a = ctx->sort_a, &b = ctx->sort_b;
  a.resize((size_t)n); b.resize((size_t)n);
  for (int i = 0; i < n; i++) {
    const uint64_t k = ((((uint64_t)(uint32_t)src[i].octave * 16 + (uint32_t)src[i].candScale) << by | (uint32_t)src[i].candY) << bx) |
                       (uint32_t)src[i].candX;
    a[i] = (k << 24) | (uint32_t)i;                       // n < 2^24 per frame; larger lists take the generic path
  }
  const int key_bits = 8 + by + bx;
  if (n >= (1 << 24) || key_bits > 40) {
    std::vector<uint32_t> idx((size_t)n);
    for (int i = 0; i < n; i++) idx[i] = (uint32_t)i;
    std::sort(idx.begin(), idx.end(), [&](uint32_t p, uint32_t q) {
      return cand_key(src[p].octave, src[p].candScale, src[p].candY, src[p].candX) <
             cand_key(src[q].octave, src[q].candScale, src[q].candY, src[q].candX);
    });
    const int m = std::min(n, dst_cap);
    for (int i = 0; i < m; i++) dst[i] = src[idx[i]];
    return;
  }
  const int DIG = 11, NB = 1 << DIG;
  uint32_t hist[1 << 11];
  for (int shift = 24; shift < 24 + key_bits; shift += DIG) {
    memset(hist, 0, sizeof hist);
    for (int i = 0; i < n; i++) hist[(a[i] >> shift) & (NB - 1)]++;
    uint32_t sum = 0;
    for (int d = 0; d < NB; d++) { const uint32_t c = hist[d]; hist[d] = sum; sum += c; }
    for (int i = 0; i < n; i++) b[hist[(a[i] >> shift) & (NB - 1)]++] = a[i];
    a.swap(b);
  }
  const int m = std::min(n, dst_cap);
  for (int i = 0; i < m; i++) dst[i] = src[a[i] & 0xffffffu];
}

static void sort_candidates(sift_candidate *c, int n)
{
  std::sort(c, c + n, [](const sift_candidate &a, const sift_candidate &b) {
    return cand_key(a.octave, a.scaleLevel, a.y, a.x) < cand_key(b.octave, b.scaleLevel, b.y, b.x);
  });
}

// Upload a host image into ctx->image (dense rows).
static int upload_image(sift_ctx *ctx, const void *image, int dtype, int w, int h, size_t pitch_bytes,
                        size_t *dev_pitch)
{
  const size_t es = dtype_size(dtype);
  if (!es) return fail(ctx, SIFT_ERR_BAD_ARGS, "unknown dtype %d", dtype);
  const size_t row = (size_t)w * es;
  if (pitch_bytes == 0) pitch_bytes = row;
  if (pitch_bytes < row) return fail(ctx, SIFT_ERR_BAD_ARGS, "pitch %zu < row bytes %zu", pitch_bytes, row);
  int rc;
  if ((rc = grow(ctx, ctx->L->image, row * h))) return rc;
  if (pitch_bytes == row) CK(cudaMemcpyAsync(ctx->L->image.p, image, row * h, cudaMemcpyHostToDevice, ctx->L->stream));
  else CK(cudaMemcpy2DAsync(ctx->L->image.p, row, image, pitch_bytes, row, h, cudaMemcpyHostToDevice, ctx->L->stream));
  ctx->bytes_h2d += row * h;
  *dev_pitch = row;
  return SIFT_OK;
}

// Download counters + keypoints (one copy in the common case), returns them sorted in h_out.
#define FIRST_CHUNK 8192
static int download_keypoints(sift_ctx *ctx, Counters *c, sift_keypoint **kps)
{
  int rc;
  const size_t first = sizeof(Counters) + (size_t)std::min(ctx->L->kp_cap, FIRST_CHUNK) * sizeof(sift_keypoint);
  if ((rc = grow_pinned(ctx, &ctx->L->h_out, &ctx->L->h_out_cap, sizeof(Counters) + (size_t)ctx->L->kp_cap * sizeof(sift_keypoint))))
    return rc;
  CK(cudaMemcpyAsync(ctx->L->h_out, ctx->L->outbuf.p, first, cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  ctx->bytes_d2h += first;
  *c = *(Counters *)ctx->L->h_out;
  const int n = std::min(c->n_kp, ctx->L->kp_cap);
  if (n > FIRST_CHUNK) ctx->bytes_d2h += (size_t)(n - FIRST_CHUNK) * sizeof(sift_keypoint);
  if (n > FIRST_CHUNK) {
    CK(cudaMemcpyAsync((char *)ctx->L->h_out + first, (char *)ctx->L->outbuf.p + first,
                       (size_t)(n - FIRST_CHUNK) * sizeof(sift_keypoint), cudaMemcpyDeviceToHost, ctx->L->stream));
    CK(cudaStreamSynchronize(ctx->L->stream));
  }
  *kps = (sift_keypoint *)((char *)ctx->L->h_out + sizeof(Counters));
  return SIFT_OK;
}

// scan + refine with automatic growth of the candidate buffer; leaves sorted keypoints in h_out.
static int scan_refine_download(sift_ctx *ctx, Counters *c, sift_keypoint **kps, int count_low)
{
  int rc;
  for (int attempt = 0; attempt < 3; attempt++) {
    CK(cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ctx->L->stream));
    if ((rc = run_scan(ctx, count_low))) return rc;
    if ((rc = run_refine(ctx, -1, dev_keypoints(ctx), ctx->L->kp_cap))) return rc;
    if ((rc = download_keypoints(ctx, c, kps))) return rc;
    if (c->n_cand <= ctx->L->cand_cap && c->n_kp <= ctx->L->kp_cap) return SIFT_OK;   // unordered: callers sort-gather
    const int want = std::max(c->n_cand, c->n_kp) + 1024;
    if ((rc = grow(ctx, ctx->L->cand, (size_t)want * sizeof(sift_candidate)))) return rc;
    ctx->L->cand_cap = want;
    if ((rc = grow(ctx, ctx->L->outbuf, sizeof(Counters) + (size_t)want * sizeof(sift_keypoint)))) return rc;
    ctx->L->kp_cap = want;
  }
  return fail(ctx, SIFT_ERR_CAPACITY, "candidate buffer kept overflowing");
}

// The device work of one batch frame on the current lane, image already in ln->image: counter reset, pyramid, scan,
// refinement into the lane's own record buffer.
static int issue_frame_kernels(sift_ctx *ctx, int dtype, size_t row_bytes)
{
  Lane *ln = ctx->L;
  int r;
  CK(cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ln->stream));
  if ((r = run_pyramid(ctx, ln->image.p, dtype, row_bytes))) return r;
  if ((r = run_scan(ctx, 0))) return r;
  return run_refine(ctx, -1, dev_keypoints(ctx), ln->kp_cap);
}

static bool same_thresholds(const sift_params &a, const sift_params &b)
{
  return same_params(a, b) && a.contrastThreshold == b.contrastThreshold && a.preFilterFactor == b.preFilterFactor &&
         a.edgeRatio == b.edgeRatio && a.maxIterations == b.maxIterations && a.offsetBound == b.offsetBound &&
         a.minInterpixelDistance == b.minInterpixelDistance;
}

// The same through a CUDA graph captured once per (plan, buffers, parameters): one launch call per frame.
static int launch_frame(sift_ctx *ctx, int dtype, size_t row_bytes)
{
  Lane *ln = ctx->L;
  if (!ctx->use_graphs || ctx->profiling || ctx->is_strip) return issue_frame_kernels(ctx, dtype, row_bytes);
  const bool valid = ln->frame_graph && ln->graph_plan == ctx->plan_id && ln->graph_bufs == ln->buf_version &&
                     ln->graph_dtype == dtype && ln->graph_keep == ctx->keep_gauss && same_thresholds(ln->graph_prm, ctx->prm);
  if (!valid) {
    if (ln->frame_graph) { cudaGraphExecDestroy(ln->frame_graph); ln->frame_graph = nullptr; }
    const int64_t l0 = ctx->launches;
    const uint64_t serial = ctx->pyramid_serial;
    if (cudaStreamBeginCapture(ln->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      (void)cudaGetLastError();
      return issue_frame_kernels(ctx, dtype, row_bytes);
    }
    const int r = issue_frame_kernels(ctx, dtype, row_bytes);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(ln->stream, &g);
    ln->graph_launches = (int)(ctx->launches - l0);
    ctx->launches = l0;
    ctx->pyramid_serial = serial;
    if (r != SIFT_OK || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      (void)cudaGetLastError();
      if (r != SIFT_OK) return r;
      return issue_frame_kernels(ctx, dtype, row_bytes);          // capture refused: plain launches
    }
    const cudaError_t ei = cudaGraphInstantiate(&ln->frame_graph, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) {
      ln->frame_graph = nullptr;
      (void)cudaGetLastError();
      return issue_frame_kernels(ctx, dtype, row_bytes);
    }
    ln->graph_plan = ctx->plan_id; ln->graph_bufs = ln->buf_version; ln->graph_dtype = dtype;
    ln->graph_keep = ctx->keep_gauss; ln->graph_prm = ctx->prm;
  }
  CK(cudaGraphLaunch(ln->frame_graph, ln->stream));
  ctx->launches += ln->graph_launches;
  ctx->pyramid_serial++;
  ctx->pyramid_built = true;
  return SIFT_OK;
}

// ------------------------------------------------------------------- C ABI ----
extern "C" {

SIFT_API const char *sift_version(void) { return SIFT_B200_VERSION; }

SIFT_API void sift_default_params(sift_params *p)
{
  if (!p) return;
  memset(p, 0, sizeof *p);
  p->numberOfOctaves = 5;          // worker.js:33
  p->scalesPerOctave = 3;          // worker.js:34
  p->minBlurLevel = 0.8;           // worker.js:35
  p->assumedBlur = 0.5;            // worker.js:36
  p->contrastThreshold = 0.015;    // sift.js:285
  p->preFilterFactor = 0.8;        // sift.js:293
  p->edgeRatio = 10.0;             // background.js:598
  p->maxIterations = 5;            // background.js:480
  p->offsetBound = 0.6;            // background.js:558
  p->minInterpixelDistance = 0.5;  // background.js:461
}

SIFT_API void sift_destroy(sift_ctx *c);
SIFT_API int sift_create(int device, sift_ctx **out)
{
  if (!out) return fail(nullptr, SIFT_ERR_BAD_ARGS, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, SIFT_ERR_NO_DEVICE, "no CUDA device (this engine has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(nullptr, SIFT_ERR_BAD_ARGS, "device %d of %d", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return fail(nullptr, SIFT_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, SIFT_ERR_NO_DEVICE, "device %d is sm_%d%d; this build is sm_100a only", device, prop.major,
                prop.minor);
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SIFT_ERR_CUDA, "cudaSetDevice failed");
  sift_ctx *c = new sift_ctx();
  c->device = device;
  bool ok = cudaStreamCreateWithFlags(&c->main_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
  for (int i = 0; ok && i < SIFT_MAX_LANES; i++)
    ok = cudaStreamCreateWithFlags(&c->lanes[i].stream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->lanes[i].ev_fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->lanes[i].ev_done, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    const char *why = cudaGetErrorString(cudaGetLastError());
    sift_destroy(c);
    return fail(nullptr, SIFT_ERR_CUDA, "stream/event creation failed: %s", why);
  }
  c->L = &c->lanes[0];
  sift_default_params(&c->prm);
  const char *fg = getenv("SIFT_B200_FORCE_GENERIC");
  c->force_generic = fg && fg[0] == '1';
  { const char *nm = getenv("SIFT_B200_NO_MMA"); c->no_mma = nm && nm[0] == '1'; }
  c->mma_big_only = getenv("SIFT_B200_MMA_BIG_ONLY") != nullptr;
  c->sep_variant_forced = getenv("SIFT_B200_NO_TMA") || getenv("SIFT_B200_NO_TMA_BLUR") || getenv("SIFT_B200_FIR_NO8");
  c->oct0_variant_forced = getenv("SIFT_B200_OCT0_WS") || getenv("SIFT_B200_OCT0_SMALL") || getenv("SIFT_B200_OCT0_BANDS") ||
                           getenv("SIFT_B200_FUSED0_LO") || getenv("SIFT_B200_FUSED0_HI1") || getenv("SIFT_B200_FUSED0_HI3");
  const char *fo = getenv("SIFT_B200_FORCE_OLD");
  c->force_old = fo && fo[0] == '1';
  const char *nt = getenv("SIFT_B200_NO_TMA");
  c->no_tma = nt && nt[0] == '1';
  const char *ng = getenv("SIFT_B200_NO_GRAPH");
  c->use_graphs = !(ng && ng[0] == '1');
  const char *nl = getenv("SIFT_B200_LANES");
  if (nl && nl[0] >= '0' && nl[0] <= '0' + SIFT_MAX_LANES) c->n_lanes = nl[0] - '0';   // 0 = by frame size
  *out = c;
  return SIFT_OK;
}

SIFT_API void sift_destroy(sift_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (Lane &ln : c->lanes) {
    Scratch *all[] = { &ln.planes, &ln.seeds, &ln.tbuf, &ln.image, &ln.cand, &ln.outbuf, &ln.tmaps, &ln.tmaps_t, &ln.tmaps_o0, &ln.tmaps_p0, &ln.order, &ln.walks };
    for (Scratch *s : all) if (s->p) cudaFree(s->p);
    if (ln.frame_graph) cudaGraphExecDestroy(ln.frame_graph);
    if (ln.d_octs) cudaFree(ln.d_octs);
    if (ln.h_out) cudaFreeHost(ln.h_out);
    if (ln.h_out2) cudaFreeHost(ln.h_out2);
    if (ln.ev_fork) cudaEventDestroy(ln.ev_fork);
    if (ln.ev_done) cudaEventDestroy(ln.ev_done);
    if (ln.stream) cudaStreamDestroy(ln.stream);
  }
  Scratch *all[] = { &c->low, &c->misc[0], &c->misc[1], &c->misc[2], &c->misc[3], &c->misc[4], &c->misc[5] };
  for (Scratch *s : all) if (s->p) cudaFree(s->p);
  if (c->d_weights) cudaFree(c->d_weights);
  if (c->h_cand) cudaFreeHost(c->h_cand);
  for (auto &sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->main_stream) cudaStreamDestroy(c->main_stream);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    for (int k = 0; k < 8; k++) cudaEventDestroy(c->strip_ev[k]);
    cudaStreamDestroy(c->copy_stream);
  }
  delete c;
}

SIFT_API const char *sift_last_error(const sift_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

// Frames in flight when the caller did not say: the grid of a small frame does not fill the GPU (a 512 x 512 frame
// is 256 octave-0 tiles for 444 resident CTAs), so more of them run side by side.  Measured on the B200
// (tools/kernel_times.py): 1080p 3 lanes = 4 lanes; 720p +6 % and 512^2 +16 % from 3 to 4 lanes, more beyond.
static int auto_lanes(const sift_ctx *ctx)
{
  if (ctx->n_lanes > 0) return ctx->n_lanes;
  const long long px = (long long)ctx->in_w * ctx->in_h;
  // 1080p: 3 = 4 = 8 lanes (6 250 Mpixel/s); 720p: 5 500 (3) -> 5 880 (6); 512^2: 3 510 (3) -> 5 210 (8);
  // 3840x2160 / 6 octaves: 5 920 (3) -> 6 080 (6)
  // with the DMMA kernels on every octave (second half of round 2): 1080p 6 860 (4) -> 7 030 (8); 720p 5 980 (4) -> 6 590 (8)
  int lanes = px >= 6000000 ? 6 : 8;
  // a lane holds a whole pyramid (~240 bytes per input pixel): keep the lanes of large images within ~32 GB
  const long long per_lane = px * 240;
  while (lanes > 1 && per_lane * lanes > (32LL << 30)) lanes--;
  return lanes;
}

// Make the public stream wait for every lane that has work in flight (non-blocking for the host).
static int join_lanes(sift_ctx *ctx)
{
  for (Lane &ln : ctx->lanes) {
    if (!ln.busy) continue;
    CK(cudaEventRecord(ln.ev_done, ln.stream));
    CK(cudaStreamWaitEvent(ctx->main_stream, ln.ev_done, 0));
    ln.busy = false;
  }
  return SIFT_OK;
}

SIFT_API int sift_flush(sift_ctx *ctx)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  CK(cudaSetDevice(ctx->device));
  return join_lanes(ctx);
}

SIFT_API int sift_synchronize(sift_ctx *ctx)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = join_lanes(ctx))) return rc;
  CK(cudaStreamSynchronize(ctx->main_stream));
  for (Lane &ln : ctx->lanes) CK(cudaStreamSynchronize(ln.stream));
  return SIFT_OK;
}

SIFT_API int sift_set_keep_gaussian(sift_ctx *ctx, int keep)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  int rc;
  if ((rc = sift_synchronize(ctx))) return rc;
  ctx->keep_gauss = keep ? 1 : 0;
  return SIFT_OK;
}

SIFT_API int sift_set_lanes(sift_ctx *ctx, int n_lanes)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (n_lanes < 0 || n_lanes > SIFT_MAX_LANES) return fail(ctx, SIFT_ERR_BAD_ARGS, "lanes %d outside 0..%d", n_lanes, SIFT_MAX_LANES);
  int rc;
  if ((rc = sift_synchronize(ctx))) return rc;
  ctx->n_lanes = n_lanes;
  ctx->next_lane = 0;
  return SIFT_OK;
}

SIFT_API void *sift_stream(sift_ctx *ctx) { return ctx ? (void *)ctx->main_stream : nullptr; }
SIFT_API int64_t sift_kernel_launches(const sift_ctx *ctx) { return ctx ? ctx->launches : 0; }
SIFT_API uint64_t sift_pyramid_serial(const sift_ctx *ctx) { return ctx ? ctx->pyramid_serial : 0; }
// Test hook (compute-sanitizer is closed on the B200 pool): overwrite every device buffer the context owns -- level
// planes, fp64 seeds, the pass intermediate, candidate / record / walk buffers -- with `byte`.  A detection whose
// result depends on what these buffers held before it ran (a read of memory the path did not write first, a tile
// that is skipped, a stale halo) then changes with the pattern; tests/test_memory_hygiene.py compares the results
// bit for bit under 0x00 / 0xFF (NaN) / 0x7F fills.
SIFT_API int sift_debug_poison(sift_ctx *ctx, int byte)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = sift_synchronize(ctx))) return rc;
  for (Lane &ln : ctx->lanes) {
    Scratch *all[] = { &ln.planes, &ln.seeds, &ln.tbuf, &ln.cand, &ln.outbuf, &ln.order, &ln.walks };
    for (Scratch *sc : all)
      if (sc->p) CK(cudaMemset(sc->p, byte, sc->cap));
  }
  ctx->pyramid_built = false;
  ctx->pyramid_serial++;
  return SIFT_OK;
}

SIFT_API void sift_transfer_bytes(const sift_ctx *ctx, uint64_t *h2d, uint64_t *d2h)
{
  if (h2d) *h2d = ctx ? ctx->bytes_h2d : 0;
  if (d2h) *d2h = ctx ? ctx->bytes_d2h : 0;
}

SIFT_API int sift_set_profiling(sift_ctx *ctx, int enabled)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  CK(cudaStreamSynchronize(ctx->L->stream));
  ctx->profiling = enabled != 0;
  ctx->spans_used = 0;
  return SIFT_OK;
}

SIFT_API int sift_get_profile(sift_ctx *ctx, float *ms_by_kind, int *launch_groups_by_kind, int n_kinds)
{
  if (!ctx || !ms_by_kind || n_kinds < SIFT_PROF_NKINDS) return SIFT_ERR_BAD_ARGS;
  CK(cudaStreamSynchronize(ctx->L->stream));
  for (int i = 0; i < n_kinds; i++) { ms_by_kind[i] = 0.f; if (launch_groups_by_kind) launch_groups_by_kind[i] = 0; }
  for (size_t i = 0; i < ctx->spans_used; i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->spans[i].a, ctx->spans[i].b);
    ms_by_kind[ctx->spans[i].kind] += ms;
    if (launch_groups_by_kind) launch_groups_by_kind[ctx->spans[i].kind]++;
  }
  ctx->spans_used = 0;
  return SIFT_OK;
}

SIFT_API int sift_detect(sift_ctx *ctx, const void *image, int dtype, int width, int height, size_t pitch_bytes,
                         const sift_params *params, sift_keypoint *out, int cap, int *n_out, sift_stats *stats)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!image || !n_out) return fail(ctx, SIFT_ERR_BAD_ARGS, "image / n_out is NULL");
  if (cap < 0 || (cap > 0 && !out)) return fail(ctx, SIFT_ERR_BAD_ARGS, "out is NULL with cap %d", cap);
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_plan0(ctx, width, height, params))) return rc;
  const int64_t l0 = ctx->launches;
  size_t dpitch;
  CK(cudaEventRecord(ctx->ev0, ctx->L->stream));
  if ((rc = upload_image(ctx, image, dtype, width, height, pitch_bytes, &dpitch))) return rc;
  if ((rc = run_pyramid(ctx, ctx->L->image.p, dtype, dpitch))) return rc;
  Counters c;
  sift_keypoint *kps;
  if ((rc = scan_refine_download(ctx, &c, &kps, 0))) return rc;
  CK(cudaEventRecord(ctx->ev1, ctx->L->stream));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->last = c;
  fill_stats(stats, c, 0, ms, (int)(ctx->launches - l0));
  *n_out = c.n_kp;
  sort_keypoints_into(ctx, kps, c.n_kp, out, cap);
  if (c.n_kp > cap) return fail(ctx, SIFT_ERR_CAPACITY, "%d keypoints, capacity %d", c.n_kp, cap);
  return SIFT_OK;
}

// *dst = number of keypoints, or -(capacity that would have been enough) when the candidate list or the output
// overflowed: which records were dropped then depends on the append order, so the caller must not use them.
__global__ void copy_count_kernel(const Counters *c, int cand_cap, int cap, int *dst)
{
  const bool overflow = c->n_cand > cand_cap || c->n_kp > cap;
  *dst = overflow ? -max(c->n_cand, c->n_kp) : c->n_kp;
}

SIFT_API int sift_detect_device(sift_ctx *ctx, const void *d_image, int dtype, int width, int height,
                                size_t pitch_bytes, const sift_params *params, sift_keypoint *d_out, int cap,
                                int *d_count, int ordered)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!d_image || !d_out || !d_count || cap <= 0) return fail(ctx, SIFT_ERR_BAD_ARGS, "NULL device pointer or cap <= 0");
  const size_t es = dtype_size(dtype);
  if (!es) return fail(ctx, SIFT_ERR_BAD_ARGS, "unknown dtype %d", dtype);
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_plan(ctx, width, height, params))) return rc;
  const int nl_eff = auto_lanes(ctx);
  if (ctx->next_lane >= nl_eff) ctx->next_lane = 0;
  Lane *ln = &ctx->lanes[ctx->next_lane];
  ctx->next_lane = (ctx->next_lane + 1) % nl_eff;
  if ((rc = ensure_lane(ctx, ln))) return rc;
  ctx->L = ln;
  if (pitch_bytes == 0) pitch_bytes = (size_t)width * es;
  // fork: everything the caller queued on the public stream so far (e.g. the producer of d_image) is visible
  CK(cudaEventRecord(ln->ev_fork, ctx->main_stream));
  CK(cudaStreamWaitEvent(ln->stream, ln->ev_fork, 0));
  rc = run_pyramid(ctx, d_image, dtype, pitch_bytes);
  if (!rc) rc = cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ln->stream) == cudaSuccess ? SIFT_OK : fail(ctx, SIFT_ERR_CUDA, "memset failed");
  if (!rc) rc = run_scan(ctx, 0);
  if (!ordered) {
    if (!rc) rc = run_refine(ctx, -1, d_out, cap);
    if (!rc) {
      copy_count_kernel<<<1, 1, 0, ln->stream>>>(dev_counters(ctx), ln->cand_cap, cap, d_count);
      ctx->launches += 1;
      if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, SIFT_ERR_CUDA, "copy_count launch failed");
    }
  } else {
    // refine into the lane's own buffer, then order (key, index) pairs on the device and gather into d_out
    const int n_sort = std::min(cap, ln->kp_cap);
    const size_t need = order_scratch_bytes(n_sort);
    if (!rc) rc = grow(ctx, ln->order, need);
    if (!rc) rc = run_refine(ctx, -1, dev_keypoints(ctx), n_sort);
    if (!rc) {
      ctx->launches += launch_order_keypoints(ln->stream, dev_keypoints(ctx), dev_counters(ctx), n_sort, ln->order.p, ln->order.cap,
                                              d_out, cap, nullptr);
      copy_count_kernel<<<1, 1, 0, ln->stream>>>(dev_counters(ctx), ln->cand_cap, n_sort, d_count);
      ctx->launches += 1;
      if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, SIFT_ERR_CUDA, "ordering launch failed");
    }
  }
  ln->busy = true;
  ctx->pyramid_built = false;          // the stage API reads lane 0, which this call may not have used
  ctx->L = &ctx->lanes[0];
  return rc;
}

static void add_stats(sift_stats &t, const sift_stats &s)
{
  t.candidates += s.candidates; t.keypoints += s.keypoints;
  t.rejLowContrast += s.rejLowContrast; t.rejEdge += s.rejEdge;
  t.rejLeftScale += s.rejLeftScale; t.rejLeftRows += s.rejLeftRows; t.rejLeftCols += s.rejLeftCols;
  t.rejNoConvergence += s.rejNoConvergence; t.rejSingular += s.rejSingular;
  t.msDevice += s.msDevice; t.kernelLaunches += s.kernelLaunches; t.leftStrip += s.leftStrip;
}

// Images are independent: frame i runs on lane i % n_lanes -- upload, kernels and download
// queued on that lane's stream -- so the copies of one frame overlap the kernels of the others and the
// host orders frame i - n_lanes + 1 while the device works.
SIFT_API int sift_detect_batch(sift_ctx *ctx, const void *images, int dtype, int width, int height,
                               size_t pitch_bytes, size_t image_stride_bytes, int n_images,
                               const sift_params *params, sift_keypoint *out, int cap, int *offsets,
                               sift_stats *stats)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!images || !offsets || n_images < 0) return fail(ctx, SIFT_ERR_BAD_ARGS, "images / offsets NULL or n_images < 0");
  if (cap < 0 || (cap > 0 && !out)) return fail(ctx, SIFT_ERR_BAD_ARGS, "out is NULL with cap %d", cap);
  const size_t es = dtype_size(dtype);
  if (!es) return fail(ctx, SIFT_ERR_BAD_ARGS, "unknown dtype %d", dtype);
  sift_stats total;
  memset(&total, 0, sizeof total);
  total.lowContrastExtrema = -1;
  offsets[0] = 0;
  if (n_images == 0) { if (stats) *stats = total; return SIFT_OK; }
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = sift_synchronize(ctx))) return rc;            // lanes may hold device-resident work of earlier calls
  if ((rc = ensure_plan(ctx, width, height, params))) return rc;
  const size_t row = (size_t)width * es;
  if (pitch_bytes == 0) pitch_bytes = row;
  if (pitch_bytes < row) return fail(ctx, SIFT_ERR_BAD_ARGS, "pitch %zu < row bytes %zu", pitch_bytes, row);
  const int NL = std::min(auto_lanes(ctx), n_images);
  auto prepare_lanes = [&]() -> int {
    for (int l = 0; l < NL; l++) {
      Lane *ln = &ctx->lanes[l];
      ctx->L = ln;
      int r;
      if ((r = ensure_lane(ctx, ln))) return r;
      if ((r = grow(ctx, ln->image, row * height))) return r;
      if ((r = grow_pinned(ctx, &ln->h_out, &ln->h_out_cap, sizeof(Counters) + (size_t)ln->kp_cap * sizeof(sift_keypoint))))
        return r;
      if ((r = grow_pinned(ctx, &ln->h_out2, &ln->h_out2_cap, sizeof(Counters) + (size_t)ln->kp_cap * sizeof(sift_keypoint))))
        return r;
    }
    return SIFT_OK;
  };
  auto hbuf = [&](Lane *ln, int frame) { return ((frame / NL) & 1) ? ln->h_out2 : ln->h_out; };
  if ((rc = prepare_lanes())) { ctx->L = &ctx->lanes[0]; return rc; }
  const int64_t l0 = ctx->launches;
  int n = 0, overflow = 0;
  static const bool trace = getenv("SIFT_B200_TRACE") != nullptr;
  double t_issue = 0, t_wait = 0, t_sort = 0;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  int done = 0, issued = 0;     // frames consumed by the host / frames whose device work has been issued
  CK(cudaEventRecord(ctx->ev0, ctx->lanes[0].stream));
  auto take = [&](int j, int ni, const sift_stats &si) { n += ni; offsets[j + 1] = n; add_stats(total, si); };
  int first_n[2 * SIFT_MAX_LANES] = { 0 };   // records covered by the first download of each frame in flight

  auto issue = [&](int i) -> int {
    const double t0 = now();
    Lane *ln = &ctx->lanes[i % NL];
    ctx->L = ln;
    const void *img = (const char *)images + (size_t)i * image_stride_bytes;
    if (pitch_bytes == row) CK(cudaMemcpyAsync(ln->image.p, img, row * height, cudaMemcpyHostToDevice, ln->stream));
    else CK(cudaMemcpy2DAsync(ln->image.p, row, img, pitch_bytes, row, height, cudaMemcpyHostToDevice, ln->stream));
    ctx->bytes_h2d += row * height;
    int r;
    if ((r = launch_frame(ctx, dtype, row))) return r;
    // counters + as many records as recent frames produced (+ 25 %): one download in the common case, without
    // shipping a fixed 8192-record block for frames that yield a thousand
    first_n[i % (2 * SIFT_MAX_LANES)] = std::min(ln->kp_cap, std::max(256, (int)(ctx->kp_running * 1.25) + 64));
    const size_t first = sizeof(Counters) + (size_t)first_n[i % (2 * SIFT_MAX_LANES)] * sizeof(sift_keypoint);
    CK(cudaMemcpyAsync(hbuf(ln, i), ln->outbuf.p, first, cudaMemcpyDeviceToHost, ln->stream));
    ctx->bytes_d2h += first;
    CK(cudaEventRecord(ln->ev_done, ln->stream));
    issued = i + 1;
    t_issue += now() - t0;
    return SIFT_OK;
  };

  // wait for frame j = `done` (device work and downloads complete).  handled: the overflow fallback consumed it.
  auto wait_frame = [&](int j, Counters &c, bool &handled) -> int {
    Lane *ln = &ctx->lanes[j % NL];
    ctx->L = ln;
    handled = false;
    const double t0 = now();
    CK(cudaEventSynchronize(ln->ev_done));
    t_wait += now() - t0;
    c = *(Counters *)hbuf(ln, j);
    if (c.n_cand > ln->cand_cap || c.n_kp > ln->kp_cap) {
      // rare: device buffers too small -- drain every lane and redo the issued frames on the growing
      // single-frame path (lane 0), then size the other lanes like it
      handled = true;
      sift_stats si;
      memset(&si, 0, sizeof si);
      for (int l = 0; l < NL; l++) CK(cudaStreamSynchronize(ctx->lanes[l].stream));
      for (int q = j; q < issued; q++) {
        int ni = 0;
        const int room = overflow ? 0 : std::max(0, cap - n);
        const void *img = (const char *)images + (size_t)q * image_stride_bytes;
        const int r2 = sift_detect(ctx, img, dtype, width, height, pitch_bytes, params, room ? out + n : nullptr, room, &ni, &si);
        if (r2 == SIFT_ERR_CAPACITY) overflow = 1;
        else if (r2 != SIFT_OK) return r2;
        take(q, ni, si);
      }
      done = issued;
      Lane *l0p = &ctx->lanes[0];
      for (int l = 1; l < NL; l++) {
        Lane *o = &ctx->lanes[l];
        ctx->L = o;
        int r;
        if (l0p->cand_cap > o->cand_cap) { if ((r = grow(ctx, o->cand, (size_t)l0p->cand_cap * sizeof(sift_candidate)))) return r; o->cand_cap = l0p->cand_cap; }
        if (l0p->kp_cap > o->kp_cap) { if ((r = grow(ctx, o->outbuf, sizeof(Counters) + (size_t)l0p->kp_cap * sizeof(sift_keypoint)))) return r; o->kp_cap = l0p->kp_cap; }
      }
      return prepare_lanes();
    }
    ctx->kp_running = std::max((double)c.n_kp, ctx->kp_running * 0.97);
    const int got = first_n[j % (2 * SIFT_MAX_LANES)];
    if (c.n_kp > got) {
      const size_t first = sizeof(Counters) + (size_t)got * sizeof(sift_keypoint);
      CK(cudaMemcpyAsync((char *)hbuf(ln, j) + first, (char *)ln->outbuf.p + first,
                         (size_t)(c.n_kp - got) * sizeof(sift_keypoint), cudaMemcpyDeviceToHost, ln->stream));
      ctx->bytes_d2h += (size_t)(c.n_kp - got) * sizeof(sift_keypoint);
      CK(cudaStreamSynchronize(ln->stream));
    }
    return SIFT_OK;
  };
  // order frame j's keypoints (in its lane's host buffer) into `out`
  auto order_frame = [&](int j, const Counters &c) {
    const double t0 = now();
    Lane *ln = &ctx->lanes[j % NL];
    sift_stats si;
    memset(&si, 0, sizeof si);
    const sift_keypoint *kps = (const sift_keypoint *)((char *)hbuf(ln, j) + sizeof(Counters));
    const int room = overflow ? 0 : std::max(0, cap - n);
    sort_keypoints_into(ctx, kps, c.n_kp, room ? out + n : nullptr, room);
    if (c.n_kp > room) overflow = 1;
    fill_stats(&si, c, 0, 0.f, 0);
    ctx->last = c;
    take(j, c.n_kp, si);
    done = j + 1;
    t_sort += now() - t0;
  };

  // Frame i runs on lane i % NL.  When frame j completes, its lane gets frame j + NL at once -- the device
  // buffers are free and the download goes to the lane's other host buffer -- and only then the host orders
  // frame j, so NL frames stay in flight while the host works.
  rc = SIFT_OK;
  while (!rc && done < n_images) {
    while (!rc && issued < n_images && issued - done < NL) rc = issue(issued);
    if (rc) break;
    const int j = done;
    Counters c;
    bool handled = false;
    rc = wait_frame(j, c, handled);
    if (rc || handled) continue;
    if (issued < n_images) rc = issue(issued);
    if (!rc) order_frame(j, c);
  }
  ctx->L = &ctx->lanes[0];
  ctx->pyramid_built = false;
  if (rc) return rc;
  for (int l = 0; l < NL; l++) CK(cudaStreamSynchronize(ctx->lanes[l].stream));
  CK(cudaEventRecord(ctx->ev1, ctx->lanes[0].stream));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  total.msDevice = ms;
  total.kernelLaunches = (int)(ctx->launches - l0);
  if (trace) fprintf(stderr, "[sift_detect_batch] %d frames: device span %.3f ms, host issue %.3f, wait %.3f, order %.3f ms\n", n_images, ms, t_issue, t_wait, t_sort);
  if (stats) *stats = total;
  if (overflow) return fail(ctx, SIFT_ERR_CAPACITY, "%d keypoints in the batch, capacity %d", n, cap);
  return SIFT_OK;
}

// ------------------------------------------------------------- stage calls ----
SIFT_API int sift_build_scale_space(sift_ctx *ctx, const void *image, int dtype, int width, int height,
                                    size_t pitch_bytes, const sift_params *params)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!image) return fail(ctx, SIFT_ERR_BAD_ARGS, "image is NULL");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_plan0(ctx, width, height, params))) return rc;
  size_t dpitch;
  if ((rc = upload_image(ctx, image, dtype, width, height, pitch_bytes, &dpitch))) return rc;
  if ((rc = run_pyramid(ctx, ctx->L->image.p, dtype, dpitch))) return rc;
  CK(cudaStreamSynchronize(ctx->L->stream));
  return SIFT_OK;
}

SIFT_API int sift_build_dog(sift_ctx *ctx)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "no scale space: call sift_build_scale_space first");
  return SIFT_OK;   // DoG levels were formed from the unrounded accumulators by the blur kernels
}

SIFT_API int sift_find_candidates(sift_ctx *ctx, const sift_params *params, sift_candidate *out, int cap,
                                  int *n_out, sift_candidate *low, int low_cap, int *n_low)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!n_out) return fail(ctx, SIFT_ERR_BAD_ARGS, "n_out is NULL");
  if (cap < 0 || low_cap < 0 || (cap > 0 && !out)) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad capacity / NULL out");
  if (!ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "no pyramid: call sift_build_scale_space first");
  CK(cudaSetDevice(ctx->device));
  const int count_low = (low != nullptr || n_low != nullptr) ? 1 : 0;
  int rc;
  Counters c;
  if (low && low_cap > ctx->low_cap) {
    if ((rc = grow(ctx, ctx->low, (size_t)low_cap * sizeof(sift_candidate)))) return rc;
    ctx->low_cap = low_cap;
  }
  for (int attempt = 0;; attempt++) {
    CK(cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ctx->L->stream));
    const int saved_low_cap = ctx->low_cap;
    if (!low) ctx->low_cap = 0;                      // count only
    rc = run_scan(ctx, count_low, params);
    ctx->low_cap = saved_low_cap;
    if (rc) return rc;
    CK(cudaMemcpyAsync(&c, dev_counters(ctx), sizeof c, cudaMemcpyDeviceToHost, ctx->L->stream));
    CK(cudaStreamSynchronize(ctx->L->stream));
    if (c.n_cand <= ctx->L->cand_cap || attempt >= 2) break;
    if ((rc = grow(ctx, ctx->L->cand, (size_t)(c.n_cand + 1024) * sizeof(sift_candidate)))) return rc;
    ctx->L->cand_cap = c.n_cand + 1024;
  }
  ctx->last = c;
  *n_out = c.n_cand;
  if (n_low) *n_low = c.n_low;
  int status = SIFT_OK;
  const int nc = std::min(std::min(c.n_cand, ctx->L->cand_cap), cap);
  if (nc > 0 && out) {
    // sort needs the whole list: stage through pinned memory
    const int all = std::min(c.n_cand, ctx->L->cand_cap);
    if ((rc = grow_pinned(ctx, &ctx->h_cand, &ctx->h_cand_cap, (size_t)all * sizeof(sift_candidate)))) return rc;
    CK(cudaMemcpyAsync(ctx->h_cand, ctx->L->cand.p, (size_t)all * sizeof(sift_candidate), cudaMemcpyDeviceToHost, ctx->L->stream));
    CK(cudaStreamSynchronize(ctx->L->stream));
    sort_candidates((sift_candidate *)ctx->h_cand, all);
    memcpy(out, ctx->h_cand, (size_t)nc * sizeof(sift_candidate));
  }
  if (c.n_cand > cap) status = fail(ctx, SIFT_ERR_CAPACITY, "%d candidates, capacity %d", c.n_cand, cap);
  if (low && c.n_low > 0) {
    const int nl = std::min(std::min(c.n_low, ctx->low_cap), low_cap);
    std::vector<sift_candidate> tmp((size_t)std::min(c.n_low, ctx->low_cap));
    CK(cudaMemcpy(tmp.data(), ctx->low.p, tmp.size() * sizeof(sift_candidate), cudaMemcpyDeviceToHost));
    sort_candidates(tmp.data(), (int)tmp.size());
    memcpy(low, tmp.data(), (size_t)nl * sizeof(sift_candidate));
    if (c.n_low > low_cap) status = fail(ctx, SIFT_ERR_CAPACITY, "%d low-contrast extrema, capacity %d", c.n_low, low_cap);
  }
  return status;
}

SIFT_API int sift_refine(sift_ctx *ctx, const sift_params *params, const sift_candidate *cands, int n_cands,
                         sift_keypoint *out, int cap, int *n_out, sift_stats *stats)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!n_out || n_cands < 0 || (n_cands > 0 && !cands)) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad candidate list");
  if (!ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "no pyramid: call sift_build_scale_space first");
  CK(cudaSetDevice(ctx->device));
  int rc;
  const int ndog = ctx->nlev - 1;
  for (int i = 0; i < n_cands; i++) {
    const sift_candidate &c = cands[i];
    if (c.octave < 0 || c.octave >= ctx->n_oct || c.scaleLevel < 1 || c.scaleLevel > ndog - 2 || c.x < 1 ||
        c.y < 1 || c.x > ctx->L->octs[c.octave].w - 2 || c.y > ctx->L->octs[c.octave].h - 2)
      return fail(ctx, SIFT_ERR_BAD_ARGS, "candidate %d (o%d s%d x%d y%d) outside the DoG interior", i, c.octave,
                  c.scaleLevel, c.x, c.y);
  }
  if (n_cands > ctx->L->cand_cap) {
    if ((rc = grow(ctx, ctx->L->cand, (size_t)n_cands * sizeof(sift_candidate)))) return rc;
    ctx->L->cand_cap = n_cands;
  }
  if (n_cands > ctx->L->kp_cap) {
    if ((rc = grow(ctx, ctx->L->outbuf, sizeof(Counters) + (size_t)n_cands * sizeof(sift_keypoint)))) return rc;
    ctx->L->kp_cap = n_cands;
  }
  CK(cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ctx->L->stream));
  Counters c;
  sift_keypoint *kps = nullptr;
  if (n_cands > 0) {
    CK(cudaMemcpyAsync(ctx->L->cand.p, cands, (size_t)n_cands * sizeof(sift_candidate), cudaMemcpyHostToDevice, ctx->L->stream));
    if ((rc = run_refine(ctx, n_cands, dev_keypoints(ctx), ctx->L->kp_cap, params))) return rc;
  }
  if ((rc = download_keypoints(ctx, &c, &kps))) return rc;
  c.n_cand = n_cands;
  // output order = order of the caller's list (background.js:468-471 iterates it as given)
  std::unordered_map<uint64_t, int> pos;
  pos.reserve((size_t)n_cands * 2);
  for (int i = n_cands - 1; i >= 0; i--) pos[cand_key(cands[i].octave, cands[i].scaleLevel, cands[i].y, cands[i].x)] = i;
  std::stable_sort(kps, kps + c.n_kp, [&](const sift_keypoint &a, const sift_keypoint &b) {
    return pos[cand_key(a.octave, a.candScale, a.candY, a.candX)] < pos[cand_key(b.octave, b.candScale, b.candY, b.candX)];
  });
  ctx->last = c;
  fill_stats(stats, c, 0, 0.f, 1);
  *n_out = c.n_kp;
  if (c.n_kp > cap) {
    if (cap > 0 && out) memcpy(out, kps, (size_t)cap * sizeof(sift_keypoint));
    return fail(ctx, SIFT_ERR_CAPACITY, "%d keypoints, capacity %d", c.n_kp, cap);
  }
  if (c.n_kp && out) memcpy(out, kps, (size_t)c.n_kp * sizeof(sift_keypoint));
  return SIFT_OK;
}

// -------------------------------------------------------- pyramid inspection ----
SIFT_API int sift_get_pyramid_info(const sift_ctx *ctx, int *octaves, int *levels_per_octave)
{
  if (!ctx || !ctx->plan_valid) return SIFT_ERR_STATE;
  if (octaves) *octaves = ctx->n_oct;
  if (levels_per_octave) *levels_per_octave = ctx->nlev;
  return SIFT_OK;
}

SIFT_API int sift_get_octave_size(const sift_ctx *ctx, int octave, int *width, int *height)
{
  if (!ctx || !ctx->plan_valid || octave < 0 || octave >= ctx->n_oct) return SIFT_ERR_BAD_ARGS;
  if (width) *width = ctx->L->octs[octave].w;
  if (height) *height = ctx->L->octs[octave].h;
  return SIFT_OK;
}

SIFT_API int sift_get_blur_level(const sift_ctx *ctx, int kind, int octave, int level, double *blur_level)
{
  if (!ctx || !ctx->plan_valid || !blur_level || octave < 0 || octave >= ctx->n_oct) return SIFT_ERR_BAD_ARGS;
  const int nl = (kind == SIFT_LEVEL_GAUSSIAN) ? ctx->nlev : ctx->nlev - 1;
  if (level < 0 || level >= nl) return SIFT_ERR_BAD_ARGS;
  *blur_level = (kind == SIFT_LEVEL_GAUSSIAN) ? ctx->plans[octave][level].blurLevel : ctx->dog_blur[octave][level];
  return SIFT_OK;
}

static float *plane_ptr(sift_ctx *ctx, int kind, int octave, int level)
{
  if (!ctx->plan_valid || octave < 0 || octave >= ctx->n_oct) return nullptr;
  const int nl = (kind == SIFT_LEVEL_GAUSSIAN) ? ctx->nlev : ctx->nlev - 1;
  if (level < 0 || level >= nl) return nullptr;
  return (kind == SIFT_LEVEL_GAUSSIAN) ? ctx->L->octs[octave].gauss[level] : ctx->L->octs[octave].dog[level];
}

SIFT_API int sift_get_level(sift_ctx *ctx, int kind, int octave, int level, float *dst)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!dst) return fail(ctx, SIFT_ERR_BAD_ARGS, "dst is NULL");
  if (!ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "no pyramid built");
  float *p = plane_ptr(ctx, kind, octave, level);
  if (!p) return fail(ctx, SIFT_ERR_BAD_ARGS, "no level kind %d octave %d level %d", kind, octave, level);
  CK(cudaSetDevice(ctx->device));
  const OctaveDev &od = ctx->L->octs[octave];
  CK(cudaMemcpy2DAsync(dst, (size_t)od.w * 4, p, (size_t)od.pitch * 4, (size_t)od.w * 4, od.h,
                       cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  return SIFT_OK;
}

SIFT_API int sift_get_level_preview(sift_ctx *ctx, int kind, int octave, int level, int mode, double coefficient,
                                    unsigned char *rgba_out, double *min_max)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!rgba_out) return fail(ctx, SIFT_ERR_BAD_ARGS, "rgba_out is NULL");
  if (mode < SIFT_PREVIEW_GRAY || mode > SIFT_PREVIEW_MINMAX) return fail(ctx, SIFT_ERR_BAD_ARGS, "unknown preview mode %d", mode);
  if (!ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "no pyramid built");
  float *p = plane_ptr(ctx, kind, octave, level);
  if (!p) return fail(ctx, SIFT_ERR_BAD_ARGS, "no level kind %d octave %d level %d", kind, octave, level);
  CK(cudaSetDevice(ctx->device));
  const OctaveDev &od = ctx->L->octs[octave];
  const size_t bytes = (size_t)od.w * od.h * 4;
  int rc;
  if ((rc = grow(ctx, ctx->L->order, bytes + 64))) return rc;          // scratch shared with the ordering path
  char *d = (char *)ctx->L->order.p;
  ctx->launches += launch_preview(ctx->L->stream, p, od.w, od.h, od.pitch, mode, coefficient, d, d + 64);
  if (cudaGetLastError() != cudaSuccess) return fail(ctx, SIFT_ERR_CUDA, "preview launch failed");
  CK(cudaMemcpyAsync(rgba_out, d + 64, bytes, cudaMemcpyDeviceToHost, ctx->L->stream));
  double mm[2] = { 0.0, 1.0 };
  CK(cudaMemcpyAsync(mm, d + 16, sizeof mm, cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  if (min_max) { min_max[0] = mm[0]; min_max[1] = mm[1]; }
  return SIFT_OK;
}

SIFT_API int sift_set_pyramid_shape(sift_ctx *ctx, int width0, int height0, const sift_params *params)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  // width0/height0 are the OCTAVE-0 dimensions (2x the input): foreign pyramids arrive already upsampled.
  if (width0 < 2 || height0 < 2 || (width0 & 1) || (height0 & 1))
    return fail(ctx, SIFT_ERR_BAD_ARGS, "octave-0 size %dx%d must be even (2x upsampled)", width0, height0);
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_plan0(ctx, width0 / 2, height0 / 2, params))) return rc;
  ctx->pyramid_built = true;
  ctx->pyramid_serial++;
  return SIFT_OK;
}

SIFT_API int sift_set_level(sift_ctx *ctx, int kind, int octave, int level, const float *src)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!src) return fail(ctx, SIFT_ERR_BAD_ARGS, "src is NULL");
  float *p = plane_ptr(ctx, kind, octave, level);
  if (!p) return fail(ctx, SIFT_ERR_BAD_ARGS, "no level kind %d octave %d level %d", kind, octave, level);
  CK(cudaSetDevice(ctx->device));
  const OctaveDev &od = ctx->L->octs[octave];
  CK(cudaMemcpy2DAsync(p, (size_t)od.pitch * 4, src, (size_t)od.w * 4, (size_t)od.w * 4, od.h,
                       cudaMemcpyHostToDevice, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  ctx->pyramid_serial++;
  return SIFT_OK;
}

// ---------------------------------------------------------- mosaic strips ----
SIFT_API int sift_strip_layout_compute(const sift_params *params, int full_width, int full_height, int row0,
                                       int row1, int margin, sift_strip_layout *out)
{
  return strip_layout(nullptr, params, full_width, full_height, row0, row1, margin, out);
}

SIFT_API int sift_strip_begin(sift_ctx *ctx, const sift_params *params, const sift_strip_layout *layout,
                              const void *rows, int dtype, size_t pitch_bytes)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!params || !layout || !rows) return fail(ctx, SIFT_ERR_BAD_ARGS, "params / layout / rows is NULL");
  if (layout->octaves != params->numberOfOctaves) return fail(ctx, SIFT_ERR_BAD_ARGS, "layout was computed for %d octaves", layout->octaves);
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = sift_synchronize(ctx))) return rc;
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));     // an upload nobody consumed (strip_begin twice)
  const int src_w = layout->width[0] / 2;
  const int src_h = (layout->bottom[0] + 1) / 2 - layout->top[0] / 2;
  ctx->L = &ctx->lanes[0];
  if ((rc = ensure_plan(ctx, src_w, src_h, params, layout))) return rc;
  if ((rc = ensure_lane(ctx, ctx->L))) return rc;
  size_t dpitch;
  ctx->strip_chunks = 0;
  const size_t es = dtype_size(dtype);
  static const bool no_overlap = getenv("SIFT_B200_STRIP_SYNC_UPLOAD") != nullptr;
  if (es && ctx->mma0_woff >= 0 && !ctx->oct0_variant_forced && !no_overlap && src_h >= 8 * 64) {
    // upload in 8 chunks of whole tile rows on a copy stream: sift_strip_octave(0) runs the octave-0 kernel band by
    // band, each band behind the chunk that holds its last halo rows, so most of the copy hides behind the blur.
    // `rows` must stay valid until sift_strip_octave(ctx, 0) has returned.
    const size_t row = (size_t)src_w * es;
    if (pitch_bytes == 0) pitch_bytes = row;
    if (pitch_bytes < row) return fail(ctx, SIFT_ERR_BAD_ARGS, "pitch %zu < row bytes %zu", pitch_bytes, row);
    if ((rc = grow(ctx, ctx->L->image, row * src_h))) return rc;
    if (!ctx->copy_stream) {
      CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 8; k++) CK(cudaEventCreateWithFlags(&ctx->strip_ev[k], cudaEventDisableTiming));
    }
    const int chunk = ((src_h + 8 * 32 - 1) / (8 * 32)) * 32;          // source rows per chunk: whole tile rows
    ctx->strip_chunk_rows = chunk;
    for (int k = 0; k < 8; k++) {
      const int r0 = std::min(k * chunk, src_h), r1 = std::min(r0 + chunk, src_h);
      if (r1 > r0)
        CK(cudaMemcpy2DAsync((char *)ctx->L->image.p + (size_t)r0 * row, row, (const char *)rows + (size_t)r0 * pitch_bytes,
                             pitch_bytes, row, r1 - r0, cudaMemcpyHostToDevice, ctx->copy_stream));
      CK(cudaEventRecord(ctx->strip_ev[k], ctx->copy_stream));
    }
    ctx->bytes_h2d += row * src_h;
    ctx->strip_chunks = 8;
    dpitch = row;
  } else {
    if ((rc = upload_image(ctx, rows, dtype, src_w, src_h, pitch_bytes, &dpitch))) return rc;
    CK(cudaStreamSynchronize(ctx->L->stream));
  }
  ctx->strip_dtype = dtype; ctx->strip_pitch = dpitch;
  ctx->strip_next_octave = 0;
  ctx->pyramid_built = false;
  ctx->pyramid_serial++;
  return SIFT_OK;
}

SIFT_API int sift_strip_seed(sift_ctx *ctx, int octave, double **d_seed)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!ctx->plan_valid || !ctx->is_strip) return fail(ctx, SIFT_ERR_STATE, "no strip: call sift_strip_begin first");
  if (!d_seed || octave < 1 || octave >= ctx->n_oct) return fail(ctx, SIFT_ERR_BAD_ARGS, "octave %d has no seed image", octave);
  *d_seed = ctx->lanes[0].octs[octave].seed64;
  return SIFT_OK;
}

SIFT_API int sift_strip_octave(sift_ctx *ctx, int octave)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!ctx->plan_valid || !ctx->is_strip) return fail(ctx, SIFT_ERR_STATE, "no strip: call sift_strip_begin first");
  if (octave != ctx->strip_next_octave || octave >= ctx->n_oct)
    return fail(ctx, SIFT_ERR_STATE, "octaves run in order: expected %d, got %d", ctx->strip_next_octave, octave);
  CK(cudaSetDevice(ctx->device));
  ctx->L = &ctx->lanes[0];
  if (octave > 0) {
    // level 0 of this octave = the seed (owned rows computed here, halo rows received): fp32 copy for read-back
    const OctaveDev &od = ctx->L->octs[octave];
    launch_seed_to_f32(ctx->L->stream, od.seed64, od.w, od.h, od.gauss[0], od.pitch);
    ctx->launches += 1;
  }
  if (octave == 0 && ctx->strip_chunks > 0) {
    // octave 0 band by band behind the chunked upload: band k needs its own rows and the first halo rows of chunk k + 1
    const OctaveDev &od = ctx->L->octs[0];
    const OctaveDev *next = ctx->n_oct > 1 ? &ctx->L->octs[1] : nullptr;
    const int per_band = ctx->strip_chunk_rows / 32, all_rows = mma0_tile_rows(od, ctx->in_h);
    prof_begin(ctx, SIFT_PROF_BLUR_OCT0);
    bool ok = true;
    for (int k = 0; k < ctx->strip_chunks && ok; k++) {
      CK(cudaStreamWaitEvent(ctx->L->stream, ctx->strip_ev[std::min(k + 1, ctx->strip_chunks - 1)], 0));
      const int r0 = k * per_band, n = (k + 1 == ctx->strip_chunks) ? all_rows - r0 : per_band;
      if (n <= 0 || r0 >= all_rows) continue;
      ok = launch_oct0_mma(ctx->L->stream, ctx->L->image.p, ctx->strip_dtype, ctx->strip_pitch, ctx->in_w, ctx->in_h, od, next,
                           ctx->d_weights + ctx->mma0_woff, ctx->plans[0], ctx->nlev, ctx->prm.scalesPerOctave, ctx->keep_gauss, r0, n);
      ctx->launches += 1;
    }
    prof_end(ctx);
    if (!ok) {                                               // layout the DMMA kernel does not take: the whole octave, after the upload
      CK(cudaStreamWaitEvent(ctx->L->stream, ctx->strip_ev[ctx->strip_chunks - 1], 0));
      run_octave(ctx, 0, ctx->L->image.p, ctx->strip_dtype, ctx->strip_pitch);
    }
    ctx->strip_chunks = 0;
  } else {
    run_octave(ctx, octave, ctx->L->image.p, ctx->strip_dtype, ctx->strip_pitch);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->L->stream));
  ctx->strip_next_octave = octave + 1;
  if (ctx->strip_next_octave == ctx->n_oct) ctx->pyramid_built = true;
  return SIFT_OK;
}

// Seed-halo exchange between the strips of ONE process (one exchange step per octave): every strip
// receives, from the strips that own them, the seed rows of `octave` it holds beyond its own range.  Device to
// device: cudaMemcpyPeerAsync on the receiving strip's stream -- over NVLink when the contexts sit on different
// GPUs, a plain device copy when they share one -- ordered after the sender's octave (octave - 1) by an event.
// The host does not wait: the receivers' next sift_strip_octave is stream-ordered after the copies.
SIFT_API int sift_mosaic_exchange(sift_ctx *const *ctxs, int n_strips, int octave)
{
  if (!ctxs || n_strips < 1) return SIFT_ERR_BAD_ARGS;
  sift_ctx *ctx = ctxs[0];
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  for (int i = 0; i < n_strips; i++) {
    sift_ctx *c = ctxs[i];
    if (!c || !c->plan_valid || !c->is_strip) return fail(ctx, SIFT_ERR_STATE, "strip %d: call sift_strip_begin first", i);
    if (octave < 1 || octave >= c->n_oct) return fail(ctx, SIFT_ERR_BAD_ARGS, "octave %d has no seed image", octave);
    if (c->strip_next_octave != octave) return fail(ctx, SIFT_ERR_STATE, "strip %d is at octave %d, not %d", i, c->strip_next_octave, octave);
    if (c->strip.width[octave] != ctx->strip.width[octave] || c->strip.height[octave] != ctx->strip.height[octave])
      return fail(ctx, SIFT_ERR_BAD_ARGS, "strip %d belongs to another mosaic", i);
  }
  // every sender's seed rows are final once its octave (octave - 1) has run: one event per strip
  for (int i = 0; i < n_strips; i++) {
    sift_ctx *c = ctxs[i];
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->lanes[0].ev_done, c->lanes[0].stream));
  }
  const size_t row_bytes = (size_t)ctx->strip.width[octave] * sizeof(double);
  for (int d = 0; d < n_strips; d++) {
    sift_ctx *dst = ctxs[d];
    const sift_strip_layout &dl = dst->strip;
    CK(cudaSetDevice(dst->device));
    for (int s2 = 0; s2 < n_strips; s2++) {             // direct NVLink path between the two GPUs (once per pair)
      int can = 0;
      if (ctxs[s2]->device != dst->device && cudaDeviceCanAccessPeer(&can, dst->device, ctxs[s2]->device) == cudaSuccess && can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[s2]->device, 0);
        if (e != cudaSuccess) (void)cudaGetLastError();   // already enabled (or refused: the copy is then staged)
      }
    }
    // halo rows = held rows outside the owned range: [top, own0) and [own1, bottom)
    const int ranges[2][2] = { { dl.top[octave], dl.own0[octave] }, { dl.own1[octave], dl.bottom[octave] } };
    for (int r = 0; r < 2; r++) {
      int row = ranges[r][0];
      while (row < ranges[r][1]) {
        int owner = -1;
        for (int s2 = 0; s2 < n_strips; s2++)
          if (row >= ctxs[s2]->strip.own0[octave] && row < ctxs[s2]->strip.own1[octave]) { owner = s2; break; }
        if (owner < 0 || owner == d) return fail(ctx, SIFT_ERR_BAD_ARGS, "row %d of octave %d is owned by no other strip", row, octave);
        sift_ctx *src = ctxs[owner];
        const int last = std::min(ranges[r][1], src->strip.own1[octave]);
        const double *sp = src->lanes[0].octs[octave].seed64 + (size_t)(row - src->strip.top[octave]) * src->strip.width[octave];
        double *dp = dst->lanes[0].octs[octave].seed64 + (size_t)(row - dl.top[octave]) * dl.width[octave];
        CK(cudaStreamWaitEvent(dst->lanes[0].stream, src->lanes[0].ev_done, 0));
        CK(cudaMemcpyPeerAsync(dp, dst->device, sp, src->device, (size_t)(last - row) * row_bytes, dst->lanes[0].stream));
        row = last;
      }
    }
  }
  // a sender must not start overwriting ... nothing: octave `octave` only reads the seed; the next octave's seed is
  // another buffer.  Receivers are stream-ordered; nothing to wait for on the host.
  return SIFT_OK;
}

static int fetch_escaped(sift_ctx *ctx, const Counters &c);
SIFT_API int sift_strip_finish(sift_ctx *ctx, sift_keypoint *out, int cap, int *n_out, sift_stats *stats)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!n_out || cap < 0 || (cap > 0 && !out)) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad output arguments");
  if (!ctx->is_strip || !ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "run every octave with sift_strip_octave first");
  CK(cudaSetDevice(ctx->device));
  ctx->L = &ctx->lanes[0];
  const int64_t l0 = ctx->launches;
  int rc;
  Counters c;
  // a strip of a gigapixel mosaic yields 10^5..10^6 records: they are ordered on the device (order.cu) and
  // downloaded once, in place, instead of being radix-sorted by the host
  for (int attempt = 0;; attempt++) {
    if (attempt == 3) return fail(ctx, SIFT_ERR_CAPACITY, "candidate buffer kept overflowing");
    CK(cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ctx->L->stream));
    if ((rc = run_scan(ctx, 0))) return rc;
    if ((rc = run_refine(ctx, -1, dev_keypoints(ctx), ctx->L->kp_cap))) return rc;
    CK(cudaMemcpyAsync(&c, dev_counters(ctx), sizeof c, cudaMemcpyDeviceToHost, ctx->L->stream));
    CK(cudaStreamSynchronize(ctx->L->stream));
    if (c.n_cand <= ctx->L->cand_cap && c.n_kp <= ctx->L->kp_cap) break;
    const int want = std::max(c.n_cand, c.n_kp) + std::max(c.n_cand, c.n_kp) / 8 + 1024;
    if ((rc = grow(ctx, ctx->L->cand, (size_t)want * sizeof(sift_candidate)))) return rc;
    ctx->L->cand_cap = want;
    if ((rc = grow(ctx, ctx->L->outbuf, sizeof(Counters) + (size_t)want * sizeof(sift_keypoint)))) return rc;
    ctx->L->kp_cap = want;
  }
  ctx->last = c;
  if ((rc = fetch_escaped(ctx, c))) return rc;
  fill_stats(stats, c, 0, 0.f, (int)(ctx->launches - l0));
  *n_out = c.n_kp;
  if (c.n_kp > cap) return fail(ctx, SIFT_ERR_CAPACITY, "%d keypoints, capacity %d", c.n_kp, cap);
  if (c.n_kp > 0) {
    const size_t sort_bytes = (order_scratch_bytes(c.n_kp) + 255) & ~(size_t)255;
    if ((rc = grow(ctx, ctx->L->order, sort_bytes + (size_t)c.n_kp * sizeof(sift_keypoint)))) return rc;
    sift_keypoint *d_sorted = (sift_keypoint *)((char *)ctx->L->order.p + sort_bytes);
    ctx->launches += launch_order_keypoints(ctx->L->stream, dev_keypoints(ctx), dev_counters(ctx), c.n_kp, ctx->L->order.p,
                                            sort_bytes, d_sorted, c.n_kp, nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_sorted, (size_t)c.n_kp * sizeof(sift_keypoint), cudaMemcpyDeviceToHost, ctx->L->stream));
    CK(cudaStreamSynchronize(ctx->L->stream));
    if (stats) stats->kernelLaunches = (int)(ctx->launches - l0);
  }
  return SIFT_OK;
}

// Download the walks the last refine launch recorded (Counters::n_left_strip of them) into ctx->escaped.
static int fetch_escaped(sift_ctx *ctx, const Counters &c)
{
  ctx->escaped.clear();
  if (c.n_left_strip <= 0) return SIFT_OK;
  if (c.n_left_strip > WALK_CAP) return fail(ctx, SIFT_ERR_CAPACITY, "%d refinement walks left the strip (capacity %d): widen the margin", c.n_left_strip, WALK_CAP);
  ctx->escaped.resize((size_t)c.n_left_strip);
  CK(cudaMemcpyAsync(ctx->escaped.data(), ctx->L->walks.p, (size_t)c.n_left_strip * sizeof(sift_walk), cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  // deterministic hand-over order, whatever order the threads finished in
  std::sort(ctx->escaped.begin(), ctx->escaped.end(), [](const sift_walk &a, const sift_walk &b) {
    return cand_key(a.octave, a.candScale, a.candY, a.candX) < cand_key(b.octave, b.candScale, b.candY, b.candX);
  });
  return SIFT_OK;
}

SIFT_API int sift_strip_escaped(sift_ctx *ctx, sift_walk *out, int cap, int *n_out)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!n_out || cap < 0 || (cap > 0 && !out)) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad output arguments");
  if (!ctx->is_strip) return fail(ctx, SIFT_ERR_STATE, "no strip in progress");
  *n_out = (int)ctx->escaped.size();
  if (*n_out > cap) return fail(ctx, SIFT_ERR_CAPACITY, "%d walks, capacity %d", *n_out, cap);
  if (*n_out) memcpy(out, ctx->escaped.data(), ctx->escaped.size() * sizeof(sift_walk));
  return SIFT_OK;
}

SIFT_API int sift_strip_resume(sift_ctx *ctx, const sift_walk *walks, int n, sift_keypoint *out, int cap, int *n_out,
                               sift_stats *stats)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!n_out || n < 0 || (n > 0 && !walks) || cap < 0 || (cap > 0 && !out)) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad arguments");
  if (!ctx->is_strip || !ctx->pyramid_built) return fail(ctx, SIFT_ERR_STATE, "run every octave with sift_strip_octave first");
  if (n > WALK_CAP) return fail(ctx, SIFT_ERR_CAPACITY, "%d walks, at most %d per call", n, WALK_CAP);
  for (int i = 0; i < n; i++) {
    const sift_walk &w = walks[i];
    if (w.octave < 0 || w.octave >= ctx->n_oct || w.scaleLevel < 1 || w.scaleLevel >= ctx->nlev - 2 || w.iteration < 0 ||
        w.x < 1 || w.x >= ctx->L->octs[w.octave].w - 1 || w.y < 1 || w.y >= ctx->L->octs[w.octave].gh - 1)
      return fail(ctx, SIFT_ERR_BAD_ARGS, "walk %d is outside the DoG volume", i);
  }
  CK(cudaSetDevice(ctx->device));
  ctx->L = &ctx->lanes[0];
  const int64_t l0 = ctx->launches;
  int rc;
  Counters c;
  memset(&c, 0, sizeof c);
  sift_keypoint *kps = nullptr;
  *n_out = 0;
  ctx->escaped.clear();
  if (n > 0) {
    if ((rc = grow(ctx, ctx->L->walks, 2 * (size_t)WALK_CAP * sizeof(sift_walk)))) return rc;
    if (n > ctx->L->kp_cap) {                                  // every walk may end as a record
      if ((rc = grow(ctx, ctx->L->outbuf, sizeof(Counters) + (size_t)n * sizeof(sift_keypoint)))) return rc;
      ctx->L->kp_cap = n;
    }
    sift_walk *d_in = (sift_walk *)ctx->L->walks.p + WALK_CAP;
    CK(cudaMemsetAsync(dev_counters(ctx), 0, sizeof(Counters), ctx->L->stream));
    CK(cudaMemcpyAsync(d_in, walks, (size_t)n * sizeof(sift_walk), cudaMemcpyHostToDevice, ctx->L->stream));
    launch_refine_resume(ctx->L->stream, ctx->L->d_octs, d_in, n, refine_params(ctx, nullptr), dev_keypoints(ctx),
                         ctx->L->kp_cap, dev_counters(ctx), (sift_walk *)ctx->L->walks.p, WALK_CAP);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if ((rc = download_keypoints(ctx, &c, &kps))) return rc;
    if ((rc = fetch_escaped(ctx, c))) return rc;
    *n_out = c.n_kp;
    sort_keypoints_into(ctx, kps, std::min(c.n_kp, ctx->L->kp_cap), out, cap);
  }
  fill_stats(stats, c, 0, 0.f, (int)(ctx->launches - l0));
  if (c.n_kp > cap) return fail(ctx, SIFT_ERR_CAPACITY, "%d keypoints, capacity %d", c.n_kp, cap);
  return SIFT_OK;
}

// ----------------------------------------------------- fine step functions ----
SIFT_API int sift_blur_chunk(sift_ctx *ctx, const double *input, int rows, int cols, double *output, double sigma,
                             int x1, int y1, int x2, int y2)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!input || !output || rows < 1 || cols < 1) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad image");
  if (x1 < 0 || y1 < 0 || x2 > cols || y2 > rows) return fail(ctx, SIFT_ERR_BAD_ARGS, "chunk outside the image");
  if (!(sigma > 0) || !std::isfinite(sigma)) return fail(ctx, SIFT_ERR_BAD_ARGS, "sigma %g", sigma);
  if (x1 >= x2 || y1 >= y2) return SIFT_OK;                      // empty chunk: the loops do not run (sift.js:96-99)
  CK(cudaSetDevice(ctx->device));
  if (!(3 * sigma <= 4096.0)) return fail(ctx, SIFT_ERR_UNSUPPORTED, "kernel radius round(3 * %g) too large (max 4096)", sigma);
  const int R = (int)js_round(3 * sigma);                        // sift.js:38
  std::vector<double> w((size_t)2 * R + 1);
  gaussian_taps(sigma, R, w.data());
  const size_t n = (size_t)rows * cols * sizeof(double);
  int rc;
  if ((rc = grow(ctx, ctx->misc[0], n)) || (rc = grow(ctx, ctx->misc[1], n)) || (rc = grow(ctx, ctx->misc[2], n)) ||
      (rc = grow(ctx, ctx->misc[3], w.size() * sizeof(double))))
    return rc;
  CK(cudaMemcpyAsync(ctx->misc[0].p, input, n, cudaMemcpyHostToDevice, ctx->L->stream));
  CK(cudaMemcpyAsync(ctx->misc[3].p, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));                        // w is a local
  launch_blur_plane_f64(ctx->L->stream, (const double *)ctx->misc[0].p, rows, cols, (double *)ctx->misc[1].p,
                        (double *)ctx->misc[2].p, (const double *)ctx->misc[3].p, R, x1, y1, x2, y2);
  ctx->launches += 2;
  CK(cudaGetLastError());
  const size_t rowb = (size_t)(x2 - x1) * sizeof(double);
  CK(cudaMemcpy2DAsync(output + (size_t)y1 * cols + x1, (size_t)cols * 8,
                       (const double *)ctx->misc[2].p + (size_t)y1 * cols + x1, (size_t)cols * 8, rowb, y2 - y1,
                       cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  return SIFT_OK;
}

SIFT_API int sift_subtract_chunk(sift_ctx *ctx, const double *a, const double *b, int rows, int cols, double *output,
                                 int x1, int y1, int x2, int y2)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!a || !b || !output || rows < 1 || cols < 1) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad image");
  if (x1 < 0 || y1 < 0 || x2 > cols || y2 > rows) return fail(ctx, SIFT_ERR_BAD_ARGS, "chunk outside the image");
  if (x1 >= x2 || y1 >= y2) return SIFT_OK;
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)rows * cols * sizeof(double);
  int rc;
  if ((rc = grow(ctx, ctx->misc[0], n)) || (rc = grow(ctx, ctx->misc[1], n)) || (rc = grow(ctx, ctx->misc[2], n))) return rc;
  CK(cudaMemcpyAsync(ctx->misc[0].p, a, n, cudaMemcpyHostToDevice, ctx->L->stream));
  CK(cudaMemcpyAsync(ctx->misc[1].p, b, n, cudaMemcpyHostToDevice, ctx->L->stream));
  launch_subtract_f64(ctx->L->stream, (const double *)ctx->misc[0].p, (const double *)ctx->misc[1].p,
                      (double *)ctx->misc[2].p, cols, x1, y1, x2, y2);
  ctx->launches += 1;
  CK(cudaGetLastError());
  const size_t rowb = (size_t)(x2 - x1) * sizeof(double);
  CK(cudaMemcpy2DAsync(output + (size_t)y1 * cols + x1, (size_t)cols * 8,
                       (const double *)ctx->misc[2].p + (size_t)y1 * cols + x1, (size_t)cols * 8, rowb, y2 - y1,
                       cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  return SIFT_OK;
}

SIFT_API int sift_find_extremas(sift_ctx *ctx, const double *d0, const double *d1, const double *d2, int rows,
                                int cols, int scales_per_octave, double contrast_threshold_c, double prefilter_factor,
                                int32_t *cand_xy, double *cand_value, int cand_cap, int *n_cand, int32_t *low_xy,
                                double *low_value, int low_cap, int *n_low)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!d0 || !d1 || !d2 || rows < 1 || cols < 1 || !n_cand || !n_low || scales_per_octave < 1)
    return fail(ctx, SIFT_ERR_BAD_ARGS, "bad arguments");
  if ((cand_cap > 0 && (!cand_xy || !cand_value)) || (low_cap > 0 && (!low_xy || !low_value)))
    return fail(ctx, SIFT_ERR_BAD_ARGS, "NULL output with non-zero capacity");
  CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)rows * cols * sizeof(double);
  const size_t npx = (size_t)rows * cols;
  int rc;
  // misc[0..2]: images; misc[3]: counts + xy + values for both lists (capacity = all pixels)
  const size_t list_bytes = npx * (2 * sizeof(int32_t) + sizeof(double));
  if ((rc = grow(ctx, ctx->misc[0], n)) || (rc = grow(ctx, ctx->misc[1], n)) || (rc = grow(ctx, ctx->misc[2], n)) ||
      (rc = grow(ctx, ctx->misc[3], 256 + 2 * list_bytes)))
    return rc;
  CK(cudaMemcpyAsync(ctx->misc[0].p, d0, n, cudaMemcpyHostToDevice, ctx->L->stream));
  CK(cudaMemcpyAsync(ctx->misc[1].p, d1, n, cudaMemcpyHostToDevice, ctx->L->stream));
  CK(cudaMemcpyAsync(ctx->misc[2].p, d2, n, cudaMemcpyHostToDevice, ctx->L->stream));
  char *base = (char *)ctx->misc[3].p;
  int *counts = (int *)base;
  double *cv = (double *)(base + 256);
  double *lv = cv + npx;
  int32_t *cxy = (int32_t *)(lv + npx);
  int32_t *lxy = cxy + 2 * npx;
  CK(cudaMemsetAsync(counts, 0, 2 * sizeof(int), ctx->L->stream));
  sift_params tmp;
  sift_default_params(&tmp);
  tmp.scalesPerOctave = scales_per_octave;
  tmp.contrastThreshold = contrast_threshold_c;
  const double pix_thr = contrast_threshold(tmp) * prefilter_factor;              // sift.js:285-293
  launch_scan_f64(ctx->L->stream, (const double *)ctx->misc[0].p, (const double *)ctx->misc[1].p,
                  (const double *)ctx->misc[2].p, rows, cols, pix_thr, cxy, cv, (int)npx, lxy, lv, (int)npx, counts);
  ctx->launches += 1;
  CK(cudaGetLastError());
  int hc[2];
  CK(cudaMemcpyAsync(hc, counts, sizeof hc, cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  *n_cand = hc[0];
  *n_low = hc[1];
  auto fetch = [&](int cnt, const int32_t *dxy, const double *dv, int32_t *oxy, double *ov, int ocap) -> int {
    if (cnt == 0 || ocap == 0) return SIFT_OK;
    std::vector<int32_t> xy((size_t)cnt * 2);
    std::vector<double> v((size_t)cnt);
    CK(cudaMemcpy(xy.data(), dxy, xy.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v.data(), dv, v.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<int> idx((size_t)cnt);
    for (int i = 0; i < cnt; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) {                           // raster order, sift.js:221-222
      return xy[2 * a + 1] != xy[2 * b + 1] ? xy[2 * a + 1] < xy[2 * b + 1] : xy[2 * a] < xy[2 * b];
    });
    const int m = std::min(cnt, ocap);
    for (int i = 0; i < m; i++) { oxy[2 * i] = xy[2 * idx[i]]; oxy[2 * i + 1] = xy[2 * idx[i] + 1]; ov[i] = v[idx[i]]; }
    return SIFT_OK;
  };
  if ((rc = fetch(hc[0], cxy, cv, cand_xy, cand_value, cand_cap))) return rc;
  if ((rc = fetch(hc[1], lxy, lv, low_xy, low_value, low_cap))) return rc;
  if (hc[0] > cand_cap || hc[1] > low_cap)
    return fail(ctx, SIFT_ERR_CAPACITY, "%d candidates / %d low-contrast, capacities %d / %d", hc[0], hc[1], cand_cap, low_cap);
  return SIFT_OK;
}

SIFT_API int sift_gradient_hessian(sift_ctx *ctx, const double *dm, const double *dc, const double *dp, int rows,
                                   int cols, int m, int n, double g[3], double h[9])
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!dm || !dc || !dp || !g || !h) return fail(ctx, SIFT_ERR_BAD_ARGS, "NULL argument");
  if (m < 1 || n < 1 || m > rows - 2 || n > cols - 2) return fail(ctx, SIFT_ERR_BAD_ARGS, "(m=%d, n=%d) has no 3x3 neighbourhood", m, n);
  CK(cudaSetDevice(ctx->device));
  // only the 3x3 neighbourhoods are needed: ship 3 rows x 3 cols of each image
  double win[3][3][3];
  const double *src[3] = { dm, dc, dp };
  for (int p = 0; p < 3; p++)
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) win[p][a][b] = src[p][(size_t)(m - 1 + a) * cols + (n - 1 + b)];
  int rc;
  if ((rc = grow(ctx, ctx->misc[4], sizeof win + 12 * sizeof(double)))) return rc;
  double *dwin = (double *)ctx->misc[4].p;
  CK(cudaMemcpyAsync(dwin, win, sizeof win, cudaMemcpyHostToDevice, ctx->L->stream));
  launch_grad_hess_f64(ctx->L->stream, dwin, dwin + 9, dwin + 18, 3, 1, 1, dwin + 27);
  ctx->launches += 1;
  CK(cudaGetLastError());
  double out12[12];
  CK(cudaMemcpyAsync(out12, dwin + 27, sizeof out12, cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  for (int i = 0; i < 3; i++) g[i] = out12[i];
  for (int i = 0; i < 9; i++) h[i] = out12[3 + i];
  return SIFT_OK;
}

SIFT_API int sift_resize_dims(int rows, int cols, double rate, int *out_rows, int *out_cols)
{
  if (rows < 0 || cols < 0 || !(rate > 0) || !std::isfinite(rate) || !out_rows || !out_cols) return SIFT_ERR_BAD_ARGS;
  // matrix2d.js:119,124 count `for (i = 0; i < rows; i += rate)`.  For a power-of-two rate every partial sum is
  // exact, so the count is ceil(rows / rate) in closed form; other rates follow the reference's accumulation,
  // bounded so that a tiny rate cannot spin (or overflow int) here.
  if ((double)std::max(rows, cols) / rate > (double)(1 << 30)) return SIFT_ERR_UNSUPPORTED;
  int e;
  if (std::frexp(rate, &e) == 0.5) {
    *out_rows = (int)std::ceil((double)rows / rate);
    *out_cols = (int)std::ceil((double)cols / rate);
    return SIFT_OK;
  }
  int r = 0, c = 0;
  for (double i = 0; i < rows; i += rate) r++;
  for (double j = 0; j < cols; j += rate) c++;
  *out_rows = r; *out_cols = c;
  return SIFT_OK;
}

SIFT_API int sift_linear_resize(sift_ctx *ctx, const double *input, int rows, int cols, double rate, double *output)
{
  if (!ctx) return SIFT_ERR_BAD_ARGS;
  if (!input || !output || rows < 1 || cols < 1 || !(rate > 0)) return fail(ctx, SIFT_ERR_BAD_ARGS, "bad arguments");
  // exact only when a*rate reproduces the reference's accumulated counter: dyadic rates (0.5, 2, 1, 4, ...)
  int e; const double mant = std::frexp(rate, &e);
  if (mant != 0.5) return fail(ctx, SIFT_ERR_UNSUPPORTED, "sampling rate %g is not a power of two", rate);
  CK(cudaSetDevice(ctx->device));
  int orows, ocols;
  if (sift_resize_dims(rows, cols, rate, &orows, &ocols) != SIFT_OK) return fail(ctx, SIFT_ERR_UNSUPPORTED, "sampling rate %g too small", rate);
  int rc;
  const size_t nin = (size_t)rows * cols * 8, nout = (size_t)orows * ocols * 8;
  if ((rc = grow(ctx, ctx->misc[0], nin)) || (rc = grow(ctx, ctx->misc[1], nout))) return rc;
  CK(cudaMemcpyAsync(ctx->misc[0].p, input, nin, cudaMemcpyHostToDevice, ctx->L->stream));
  launch_resize_f64(ctx->L->stream, (const double *)ctx->misc[0].p, rows, cols, rate, (double *)ctx->misc[1].p, orows, ocols);
  ctx->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(output, ctx->misc[1].p, nout, cudaMemcpyDeviceToHost, ctx->L->stream));
  CK(cudaStreamSynchronize(ctx->L->stream));
  return SIFT_OK;
}

}  // extern "C"
