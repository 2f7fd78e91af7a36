/*
 * sift_addon.c -- thin Node.js N-API addon over the C ABI of include/sift_b200.h.
 *
 * This is the binding BASELINE.json's north_star asks for ("JS host code (Node.js) calls CUDA through a
 * thin C-ABI N-API addon"): it replaces the reference's worker message path (src/worker.js:29-98 senders,
 * background.js:14-50 onmessage switch).  It owns no algorithm: every function unpacks typed arrays /
 * plain numbers, calls one sift_* entry point and packs the result.  Matrix2D <-> typed-array conversion
 * and the reference's reply schemas live in addon/sift.js and addon/background.js.
 *
 * Build (needs only a C compiler; N-API symbols resolve from the node executable at load time):
 *   gcc -O2 -fPIC -shared -I../include sift_addon.c -L<dir of libsift_b200.so> -lsift_b200 \
 *       -Wl,-rpath,'$ORIGIN' -o sift_b200.node
 *
 * Records cross the boundary as raw bytes (ArrayBuffer): 80-byte sift_keypoint and 24-byte sift_candidate
 * exactly as laid out in sift_b200.h; the JS side decodes them with a DataView.
 */
#ifdef SIFT_ADDON_USE_SYSTEM_NODE_API
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sift_b200.h"

#define MAX_ARGS 12
#define ARGS(n)                                                          \
  size_t argc = (n);                                                     \
  napi_value argv[MAX_ARGS];                                             \
  if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < (n)) \
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "too few arguments")

static napi_value throw_msg(napi_env env, const char *code, const char *msg)
{
  napi_throw_error(env, code, msg);
  return NULL;
}

static const char *status_name(int rc)
{
  switch (rc) {
    case SIFT_ERR_BAD_ARGS: return "SIFT_ERR_BAD_ARGS";
    case SIFT_ERR_CUDA: return "SIFT_ERR_CUDA";
    case SIFT_ERR_CAPACITY: return "SIFT_ERR_CAPACITY";
    case SIFT_ERR_UNSUPPORTED: return "SIFT_ERR_UNSUPPORTED";
    case SIFT_ERR_NO_DEVICE: return "SIFT_ERR_NO_DEVICE";
    case SIFT_ERR_STATE: return "SIFT_ERR_STATE";
  }
  return "SIFT_ERR";
}

/* non-zero status -> thrown JS Error carrying sift_last_error() (SURVEY.md 8b error convention) */
static napi_value throw_status(napi_env env, sift_ctx *ctx, int rc)
{
  napi_throw_error(env, status_name(rc), sift_last_error(ctx));
  return NULL;
}

static sift_ctx *get_ctx(napi_env env, napi_value v)
{
  void *p = NULL;
  if (napi_get_value_external(env, v, &p) != napi_ok) return NULL;
  return (sift_ctx *)p;
}

static int get_i32(napi_env env, napi_value v) { int32_t x = 0; napi_get_value_int32(env, v, &x); return x; }
static double get_f64(napi_env env, napi_value v) { double x = 0; napi_get_value_double(env, v, &x); return x; }

static void set_num(napi_env env, napi_value obj, const char *name, double v)
{
  napi_value n;
  napi_create_double(env, v, &n);
  napi_set_named_property(env, obj, name, n);
}

/* typed array / ArrayBuffer -> pointer + byte length */
static void *get_bytes(napi_env env, napi_value v, size_t *bytes)
{
  bool is = false;
  void *data = NULL;
  if (napi_is_typedarray(env, v, &is) == napi_ok && is) {
    napi_typedarray_type t;
    size_t len = 0, off = 0;
    napi_value ab;
    if (napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) != napi_ok) return NULL;
    static const size_t es[] = { 1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8 };     /* napi_int8_array .. napi_biguint64_array */
    if ((size_t)t >= sizeof es / sizeof es[0]) return NULL;             /* element types newer than this table */
    *bytes = len * es[t];
    return data;
  }
  if (napi_is_arraybuffer(env, v, &is) == napi_ok && is) {
    if (napi_get_arraybuffer_info(env, v, &data, bytes) != napi_ok) return NULL;
    return data;
  }
  return NULL;
}

/* {numberOfOctaves, scalesPerOctave, minBlurLevel, assumedBlur, ...}: the reference's request fields
 * (worker.js:40-48, 90-98); missing fields keep the reference defaults (sift_default_params). */
static void get_params(napi_env env, napi_value obj, sift_params *p)
{
  sift_default_params(p);
  napi_valuetype t;
  if (napi_typeof(env, obj, &t) != napi_ok || t != napi_object) return;
#define FIELD(name, conv)                                                         \
  do {                                                                            \
    bool has = false;                                                             \
    napi_value v;                                                                 \
    if (napi_has_named_property(env, obj, #name, &has) == napi_ok && has &&       \
        napi_get_named_property(env, obj, #name, &v) == napi_ok &&                \
        napi_typeof(env, v, &t) == napi_ok && t == napi_number)                   \
      p->name = conv(env, v);                                                     \
  } while (0)
  FIELD(numberOfOctaves, get_i32); FIELD(scalesPerOctave, get_i32); FIELD(minBlurLevel, get_f64);
  FIELD(assumedBlur, get_f64); FIELD(contrastThreshold, get_f64); FIELD(preFilterFactor, get_f64);
  FIELD(edgeRatio, get_f64); FIELD(maxIterations, get_i32); FIELD(offsetBound, get_f64);
  FIELD(minInterpixelDistance, get_f64);
#undef FIELD
}

static napi_value stats_object(napi_env env, const sift_stats *s)
{
  napi_value o;
  napi_create_object(env, &o);
  set_num(env, o, "candidates", s->candidates); set_num(env, o, "lowContrastExtrema", s->lowContrastExtrema);
  set_num(env, o, "keypoints", s->keypoints); set_num(env, o, "rejLowContrast", s->rejLowContrast);
  set_num(env, o, "rejEdge", s->rejEdge); set_num(env, o, "rejLeftScale", s->rejLeftScale);
  set_num(env, o, "rejLeftRows", s->rejLeftRows); set_num(env, o, "rejLeftCols", s->rejLeftCols);
  set_num(env, o, "rejNoConvergence", s->rejNoConvergence); set_num(env, o, "rejSingular", s->rejSingular);
  set_num(env, o, "msDevice", s->msDevice); set_num(env, o, "kernelLaunches", s->kernelLaunches);
  set_num(env, o, "leftStrip", s->leftStrip);
  return o;
}

static void finalize_ctx(napi_env env, void *data, void *hint)
{
  (void)env; (void)hint;
  sift_destroy((sift_ctx *)data);
}

/* create(device) -> context handle (destroyed by the garbage collector or destroy()) */
static napi_value Create(napi_env env, napi_callback_info info)
{
  ARGS(1);
  sift_ctx *ctx = NULL;
  const int rc = sift_create(get_i32(env, argv[0]), &ctx);
  if (rc != SIFT_OK) return throw_status(env, NULL, rc);          /* no CPU fallback: creation fails loudly */
  napi_value ext;
  napi_create_external(env, ctx, finalize_ctx, NULL, &ext);
  return ext;
}

static napi_value Version(napi_env env, napi_callback_info info)
{
  (void)info;
  napi_value s;
  napi_create_string_utf8(env, sift_version(), NAPI_AUTO_LENGTH, &s);
  return s;
}

/* detect(ctx, pixels, width, height, dtype, params[, capacity]) -> {count, records:ArrayBuffer, stats} */
static napi_value Detect(napi_env env, napi_callback_info info)
{
  ARGS(6);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t bytes = 0;
  void *px = get_bytes(env, argv[1], &bytes);
  const int w = get_i32(env, argv[2]), h = get_i32(env, argv[3]), dtype = get_i32(env, argv[4]);
  if (!ctx || !px) return throw_msg(env, "SIFT_ERR_BAD_ARGS", "context / pixel buffer expected");
  static const size_t es[] = { 1, 4, 8, 4 };
  if (dtype < 0 || dtype > 3 || w < 1 || h < 1 || bytes < (size_t)w * h * es[dtype])
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "pixel buffer smaller than width * height");
  sift_params prm;
  get_params(env, argv[5], &prm);
  int cap = argc > 6 ? get_i32(env, argv[6]) : 0;
  if (cap <= 0) cap = (w * h) / 64 > 4096 ? (w * h) / 64 : 4096;
  for (;;) {
    void *rec = NULL;
    napi_value ab;
    if (napi_create_arraybuffer(env, (size_t)cap * sizeof(sift_keypoint), &rec, &ab) != napi_ok) return NULL;
    int n = 0;
    sift_stats st;
    const int rc = sift_detect(ctx, px, dtype, w, h, 0, &prm, (sift_keypoint *)rec, cap, &n, &st);
    if (rc == SIFT_ERR_CAPACITY && n > cap) { cap = n; continue; }  /* n holds the required count; any other capacity
                                                                     * failure (device buffers kept overflowing) throws */
    if (rc != SIFT_OK) return throw_status(env, ctx, rc);
    napi_value out;
    napi_create_object(env, &out);
    set_num(env, out, "count", n);
    napi_set_named_property(env, out, "records", ab);
    napi_set_named_property(env, out, "stats", stats_object(env, &st));
    return out;
  }
}

/* detectBatch(ctx, pixels, width, height, nImages, dtype, params, capacity) -> {records, offsets:Int32Array, stats} */
static napi_value DetectBatch(napi_env env, napi_callback_info info)
{
  ARGS(8);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t bytes = 0;
  void *px = get_bytes(env, argv[1], &bytes);
  const int w = get_i32(env, argv[2]), h = get_i32(env, argv[3]), n_img = get_i32(env, argv[4]), dtype = get_i32(env, argv[5]);
  static const size_t es[] = { 1, 4, 8, 4 };
  if (!ctx || !px || dtype < 0 || dtype > 3 || n_img < 0 || bytes < (size_t)w * h * es[dtype] * (size_t)n_img)
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "context / pixel buffer of nImages * width * height expected");
  sift_params prm;
  get_params(env, argv[6], &prm);
  const int cap = get_i32(env, argv[7]);
  void *rec = NULL, *offs = NULL;
  napi_value ab, oab, oarr;
  napi_create_arraybuffer(env, (size_t)(cap > 0 ? cap : 1) * sizeof(sift_keypoint), &rec, &ab);
  napi_create_arraybuffer(env, (size_t)(n_img + 1) * sizeof(int32_t), &offs, &oab);
  napi_create_typedarray(env, napi_int32_array, (size_t)n_img + 1, oab, 0, &oarr);
  sift_stats st;
  const int rc = sift_detect_batch(ctx, px, dtype, w, h, 0, (size_t)w * h * es[dtype], n_img, &prm, (sift_keypoint *)rec, cap,
                                   (int *)offs, &st);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_value out;
  napi_create_object(env, &out);
  napi_set_named_property(env, out, "records", ab);
  napi_set_named_property(env, out, "offsets", oarr);
  napi_set_named_property(env, out, "stats", stats_object(env, &st));
  return out;
}

/* buildScaleSpace(ctx, pixels, width, height, dtype, params) -- background.js:71 */
static napi_value BuildScaleSpace(napi_env env, napi_callback_info info)
{
  ARGS(6);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t bytes = 0;
  void *px = get_bytes(env, argv[1], &bytes);
  if (!ctx || !px) return throw_msg(env, "SIFT_ERR_BAD_ARGS", "context / pixel buffer expected");
  const int w = get_i32(env, argv[2]), h = get_i32(env, argv[3]), dtype = get_i32(env, argv[4]);
  static const size_t px_bytes[] = { 1, 4, 8, 4 };                      /* SIFT_U8, SIFT_F32, SIFT_F64, SIFT_RGBA8 */
  if (w < 1 || h < 1 || dtype < 0 || dtype > 3 || bytes < (size_t)w * h * px_bytes[dtype])
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "pixel buffer smaller than width * height of the given dtype");
  sift_params prm;
  get_params(env, argv[5], &prm);
  const int rc = sift_build_scale_space(ctx, px, dtype, w, h, 0, &prm);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  return NULL;
}

/* pyramidSerial(ctx) -> generation of the pyramid the stage calls read (sift_pyramid_serial) */
static napi_value PyramidSerial(napi_env env, napi_callback_info info)
{
  ARGS(1);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  if (!ctx) return throw_msg(env, "SIFT_ERR_BAD_ARGS", "context expected");
  napi_value n;
  napi_create_double(env, (double)sift_pyramid_serial(ctx), &n);
  return n;
}

/* pyramidInfo(ctx) -> {octaves, levels, sizes:[w0,h0,w1,h1,...] as Int32Array} */
static napi_value PyramidInfo(napi_env env, napi_callback_info info)
{
  ARGS(1);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  int o = 0, l = 0;
  const int rc = sift_get_pyramid_info(ctx, &o, &l);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  void *p = NULL;
  napi_value ab, arr, out;
  napi_create_arraybuffer(env, (size_t)o * 2 * sizeof(int32_t), &p, &ab);
  for (int i = 0; i < o; i++) sift_get_octave_size(ctx, i, (int *)p + 2 * i, (int *)p + 2 * i + 1);
  napi_create_typedarray(env, napi_int32_array, (size_t)o * 2, ab, 0, &arr);
  napi_create_object(env, &out);
  set_num(env, out, "octaves", o);
  set_num(env, out, "levels", l);
  napi_set_named_property(env, out, "sizes", arr);
  return out;
}

/* getLevel(ctx, kind, octave, level) -> {blurLevel, width, height, data:Float32Array} */
static napi_value GetLevel(napi_env env, napi_callback_info info)
{
  ARGS(4);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  const int kind = get_i32(env, argv[1]), o = get_i32(env, argv[2]), s = get_i32(env, argv[3]);
  int w = 0, h = 0;
  double blur = 0;
  int rc = sift_get_octave_size(ctx, o, &w, &h);
  if (rc == SIFT_OK) rc = sift_get_blur_level(ctx, kind, o, s, &blur);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  void *p = NULL;
  napi_value ab, arr, out;
  napi_create_arraybuffer(env, (size_t)w * h * sizeof(float), &p, &ab);
  rc = sift_get_level(ctx, kind, o, s, (float *)p);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_create_typedarray(env, napi_float32_array, (size_t)w * h, ab, 0, &arr);
  napi_create_object(env, &out);
  set_num(env, out, "blurLevel", blur);
  set_num(env, out, "width", w);
  set_num(env, out, "height", h);
  napi_set_named_property(env, out, "data", arr);
  return out;
}

/* levelPreview(ctx, kind, octave, level, mode, coefficient) -> {width, height, data:Uint8ClampedArray, min, max}
 * -- the ImageData payload the reference posts beside a stage result (image-utils.js:171-220; mode 1 =
 * Matrix2D_sigmoidNormalize first, mode 2 = Matrix2D_sampledNormalize first) */
static napi_value LevelPreview(napi_env env, napi_callback_info info)
{
  ARGS(6);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  const int kind = get_i32(env, argv[1]), o = get_i32(env, argv[2]), s = get_i32(env, argv[3]);
  const int mode = get_i32(env, argv[4]);
  double coefficient = 1.0, mm[2] = { 0.0, 1.0 };
  napi_get_value_double(env, argv[5], &coefficient);
  int w = 0, h = 0;
  int rc = sift_get_octave_size(ctx, o, &w, &h);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  void *p = NULL;
  napi_value ab, arr, out;
  napi_create_arraybuffer(env, (size_t)w * h * 4, &p, &ab);
  rc = sift_get_level_preview(ctx, kind, o, s, mode, coefficient, (unsigned char *)p, mm);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_create_typedarray(env, napi_uint8_clamped_array, (size_t)w * h * 4, ab, 0, &arr);
  napi_create_object(env, &out);
  set_num(env, out, "width", w);
  set_num(env, out, "height", h);
  set_num(env, out, "min", mm[0]);
  set_num(env, out, "max", mm[1]);
  napi_set_named_property(env, out, "data", arr);
  return out;
}

/* setPyramidShape(ctx, width0, height0, params); setLevel(ctx, kind, octave, level, Float32Array) */
static napi_value SetPyramidShape(napi_env env, napi_callback_info info)
{
  ARGS(4);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  sift_params prm;
  get_params(env, argv[3], &prm);
  const int rc = sift_set_pyramid_shape(ctx, get_i32(env, argv[1]), get_i32(env, argv[2]), &prm);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  return NULL;
}

static napi_value SetLevel(napi_env env, napi_callback_info info)
{
  ARGS(5);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t bytes = 0;
  void *p = get_bytes(env, argv[4], &bytes);
  const int o = get_i32(env, argv[2]);
  int w = 0, h = 0;
  if (!p || sift_get_octave_size(ctx, o, &w, &h) != SIFT_OK || bytes < (size_t)w * h * sizeof(float))
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "Float32Array of the octave's size expected");
  const int rc = sift_set_level(ctx, get_i32(env, argv[1]), o, get_i32(env, argv[3]), (const float *)p);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  return NULL;
}

/* findCandidates(ctx, params|null, wantLow) -> {count, records, lowCount, lowRecords} -- background.js:359 */
static napi_value FindCandidates(napi_env env, napi_callback_info info)
{
  ARGS(3);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  napi_valuetype t;
  sift_params prm, *pp = NULL;
  if (napi_typeof(env, argv[1], &t) == napi_ok && t == napi_object) { get_params(env, argv[1], &prm); pp = &prm; }
  bool want_low = false;
  napi_get_value_bool(env, argv[2], &want_low);
  int cap = 1 << 16;
  for (;;) {
    void *rec = NULL, *lrec = NULL;
    napi_value ab, lab;
    napi_create_arraybuffer(env, (size_t)cap * sizeof(sift_candidate), &rec, &ab);
    napi_create_arraybuffer(env, (size_t)(want_low ? cap : 1) * sizeof(sift_candidate), &lrec, &lab);
    int n = 0, nl = 0;
    const int rc = sift_find_candidates(ctx, pp, (sift_candidate *)rec, cap, &n, want_low ? (sift_candidate *)lrec : NULL,
                                        want_low ? cap : 0, want_low ? &nl : NULL);
    if (rc == SIFT_ERR_CAPACITY && (n > cap || nl > cap)) { cap = (n > nl ? n : nl); continue; }
    if (rc != SIFT_OK) return throw_status(env, ctx, rc);
    napi_value out;
    napi_create_object(env, &out);
    set_num(env, out, "count", n);
    napi_set_named_property(env, out, "records", ab);
    set_num(env, out, "lowCount", nl);
    napi_set_named_property(env, out, "lowRecords", lab);
    return out;
  }
}

/* refine(ctx, params, candidateRecords:ArrayBuffer, nCandidates) -> {count, records, stats} -- background.js:455 */
static napi_value Refine(napi_env env, napi_callback_info info)
{
  ARGS(4);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  sift_params prm;
  get_params(env, argv[1], &prm);
  size_t bytes = 0;
  void *cands = get_bytes(env, argv[2], &bytes);
  const int n_c = get_i32(env, argv[3]);
  if (n_c < 0 || (n_c > 0 && (!cands || bytes < (size_t)n_c * sizeof(sift_candidate))))
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "candidate records expected");
  void *rec = NULL;
  napi_value ab;
  napi_create_arraybuffer(env, (size_t)(n_c > 0 ? n_c : 1) * sizeof(sift_keypoint), &rec, &ab);
  int n = 0;
  sift_stats st;
  const int rc = sift_refine(ctx, &prm, (const sift_candidate *)cands, n_c, (sift_keypoint *)rec, n_c > 0 ? n_c : 1, &n, &st);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_value out;
  napi_create_object(env, &out);
  set_num(env, out, "count", n);
  napi_set_named_property(env, out, "records", ab);
  napi_set_named_property(env, out, "stats", stats_object(env, &st));
  return out;
}

/* ---- the five src/sift.js step functions on Float64Array images (row-major, rows x cols) ---- */
/* blurChunk(ctx, input, rows, cols, output, sigma, x1, y1, x2, y2) -- sift.js:72: writes the chunk of output */
static napi_value BlurChunk(napi_env env, napi_callback_info info)
{
  ARGS(10);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t bi = 0, bo = 0;
  double *in = (double *)get_bytes(env, argv[1], &bi), *out = (double *)get_bytes(env, argv[4], &bo);
  const int rows = get_i32(env, argv[2]), cols = get_i32(env, argv[3]);
  if (!in || !out || bi < (size_t)rows * cols * 8 || bo < (size_t)rows * cols * 8)
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "Float64Array images of rows * cols expected");
  const int rc = sift_blur_chunk(ctx, in, rows, cols, out, get_f64(env, argv[5]), get_i32(env, argv[6]), get_i32(env, argv[7]),
                                 get_i32(env, argv[8]), get_i32(env, argv[9]));
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  return NULL;
}

/* subtractChunk(ctx, a, b, rows, cols, output, x1, y1, x2, y2) -- sift.js:154 */
static napi_value SubtractChunk(napi_env env, napi_callback_info info)
{
  ARGS(10);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t ba = 0, bb = 0, bo = 0;
  double *a = (double *)get_bytes(env, argv[1], &ba), *b = (double *)get_bytes(env, argv[2], &bb);
  double *out = (double *)get_bytes(env, argv[5], &bo);
  const int rows = get_i32(env, argv[3]), cols = get_i32(env, argv[4]);
  const size_t need = (size_t)rows * cols * 8;
  if (!a || !b || !out || ba < need || bb < need || bo < need)
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "Float64Array images of rows * cols expected");
  const int rc = sift_subtract_chunk(ctx, a, b, rows, cols, out, get_i32(env, argv[6]), get_i32(env, argv[7]),
                                     get_i32(env, argv[8]), get_i32(env, argv[9]));
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  return NULL;
}

/* findExtremas(ctx, d0, d1, d2, rows, cols, scalesPerOctave, contrast, preFilter)
 *   -> {nCand, candXY:Int32Array, candValue:Float64Array, nLow, lowXY, lowValue} -- sift.js:212 */
static napi_value FindExtremas(napi_env env, napi_callback_info info)
{
  ARGS(9);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t b0 = 0, b1 = 0, b2 = 0;
  double *d0 = (double *)get_bytes(env, argv[1], &b0), *d1 = (double *)get_bytes(env, argv[2], &b1);
  double *d2 = (double *)get_bytes(env, argv[3], &b2);
  const int rows = get_i32(env, argv[4]), cols = get_i32(env, argv[5]);
  const size_t npx = (size_t)rows * cols;
  if (!d0 || !d1 || !d2 || b0 < npx * 8 || b1 < npx * 8 || b2 < npx * 8)
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "three Float64Array images of rows * cols expected");
  void *cxy, *cv, *lxy, *lv;
  napi_value acxy, acv, alxy, alv, tcxy, tcv, tlxy, tlv;
  napi_create_arraybuffer(env, npx * 2 * sizeof(int32_t), &cxy, &acxy);
  napi_create_arraybuffer(env, npx * sizeof(double), &cv, &acv);
  napi_create_arraybuffer(env, npx * 2 * sizeof(int32_t), &lxy, &alxy);
  napi_create_arraybuffer(env, npx * sizeof(double), &lv, &alv);
  int nc = 0, nl = 0;
  const int rc = sift_find_extremas(ctx, d0, d1, d2, rows, cols, get_i32(env, argv[6]), get_f64(env, argv[7]), get_f64(env, argv[8]),
                                    (int32_t *)cxy, (double *)cv, (int)npx, &nc, (int32_t *)lxy, (double *)lv, (int)npx, &nl);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_create_typedarray(env, napi_int32_array, (size_t)nc * 2, acxy, 0, &tcxy);
  napi_create_typedarray(env, napi_float64_array, (size_t)nc, acv, 0, &tcv);
  napi_create_typedarray(env, napi_int32_array, (size_t)nl * 2, alxy, 0, &tlxy);
  napi_create_typedarray(env, napi_float64_array, (size_t)nl, alv, 0, &tlv);
  napi_value out;
  napi_create_object(env, &out);
  set_num(env, out, "nCand", nc);
  napi_set_named_property(env, out, "candXY", tcxy);
  napi_set_named_property(env, out, "candValue", tcv);
  set_num(env, out, "nLow", nl);
  napi_set_named_property(env, out, "lowXY", tlxy);
  napi_set_named_property(env, out, "lowValue", tlv);
  return out;
}

/* gradientHessian(ctx, dm, dc, dp, rows, cols, m, n) -> Float64Array(12): g[3] then h[9] -- sift.js:333, 377 */
static napi_value GradientHessian(napi_env env, napi_callback_info info)
{
  ARGS(8);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t b0 = 0, b1 = 0, b2 = 0;
  double *dm = (double *)get_bytes(env, argv[1], &b0), *dc = (double *)get_bytes(env, argv[2], &b1);
  double *dp = (double *)get_bytes(env, argv[3], &b2);
  const int rows = get_i32(env, argv[4]), cols = get_i32(env, argv[5]);
  const size_t need = (size_t)rows * cols * 8;
  if (!dm || !dc || !dp || b0 < need || b1 < need || b2 < need)
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "three Float64Array images of rows * cols expected");
  void *p = NULL;
  napi_value ab, arr;
  napi_create_arraybuffer(env, 12 * sizeof(double), &p, &ab);
  const int rc = sift_gradient_hessian(ctx, dm, dc, dp, rows, cols, get_i32(env, argv[6]), get_i32(env, argv[7]), (double *)p,
                                       (double *)p + 3);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_create_typedarray(env, napi_float64_array, 12, ab, 0, &arr);
  return arr;
}

/* linearResize(ctx, input, rows, cols, rate) -> {rows, cols, data:Float64Array} -- matrix2d.js:112 */
static napi_value LinearResize(napi_env env, napi_callback_info info)
{
  ARGS(5);
  sift_ctx *ctx = get_ctx(env, argv[0]);
  size_t bi = 0;
  double *in = (double *)get_bytes(env, argv[1], &bi);
  const int rows = get_i32(env, argv[2]), cols = get_i32(env, argv[3]);
  const double rate = get_f64(env, argv[4]);
  int orows = 0, ocols = 0;
  if (!in || bi < (size_t)rows * cols * 8 || sift_resize_dims(rows, cols, rate, &orows, &ocols) != SIFT_OK)
    return throw_msg(env, "SIFT_ERR_BAD_ARGS", "Float64Array image of rows * cols and a positive rate expected");
  void *p = NULL;
  napi_value ab, arr, out;
  napi_create_arraybuffer(env, (size_t)orows * ocols * sizeof(double), &p, &ab);
  const int rc = sift_linear_resize(ctx, in, rows, cols, rate, (double *)p);
  if (rc != SIFT_OK) return throw_status(env, ctx, rc);
  napi_create_typedarray(env, napi_float64_array, (size_t)orows * ocols, ab, 0, &arr);
  napi_create_object(env, &out);
  set_num(env, out, "rows", orows);
  set_num(env, out, "cols", ocols);
  napi_set_named_property(env, out, "data", arr);
  return out;
}

NAPI_EXTERN napi_value napi_register_module_v1(napi_env env, napi_value exports)
{
  static const struct { const char *name; napi_callback fn; } table[] = {
    { "create", Create }, { "version", Version }, { "detect", Detect }, { "detectBatch", DetectBatch },
    { "buildScaleSpace", BuildScaleSpace }, { "pyramidInfo", PyramidInfo }, { "pyramidSerial", PyramidSerial }, { "getLevel", GetLevel }, { "levelPreview", LevelPreview },
    { "setPyramidShape", SetPyramidShape }, { "setLevel", SetLevel }, { "findCandidates", FindCandidates },
    { "refine", Refine }, { "blurChunk", BlurChunk }, { "subtractChunk", SubtractChunk },
    { "findExtremas", FindExtremas }, { "gradientHessian", GradientHessian }, { "linearResize", LinearResize },
  };
  for (size_t i = 0; i < sizeof table / sizeof table[0]; i++) {
    napi_property_descriptor d = { table[i].name, NULL, table[i].fn, NULL, NULL, NULL, napi_enumerable, NULL };
    if (napi_define_properties(env, exports, 1, &d) != napi_ok) return NULL;
  }
  return exports;
}
