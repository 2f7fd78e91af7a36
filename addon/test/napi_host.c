/*
 * napi_host.c -- a minimal Node-API HOST for tests (this image has no Node.js).
 *
 * It implements the ~25 napi_* functions addon/sift_addon.c uses with a tiny value model, exports them from
 * the executable (-rdynamic) exactly like the `node` binary does, dlopen()s sift_b200.node, calls its
 * napi_register_module_v1 and then drives the exported functions the way addon/background.js does.
 * Not a JS engine: it only proves that the addon registers, unpacks typed arrays / option objects, calls
 * the C ABI and packs results and errors correctly.
 *
 *   napi_host <sift_b200.node> version
 *   napi_host <sift_b200.node> create                      -> "created" or "threw <code>: <message>"
 *   napi_host <sift_b200.node> detect <raw u8 file> <w> <h> <octaves> <minBlur> <out records file>
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../node_api_min.h"

typedef enum { V_UNDEF, V_NUM, V_BOOL, V_STR, V_OBJ, V_FUNC, V_EXT, V_AB, V_TA } vkind;
typedef struct prop { char *name; struct napi_value__ *v; struct prop *next; } prop;
struct napi_value__ {
  vkind kind;
  double num;
  char *str;
  void *ptr;            /* external / arraybuffer data / typed array data */
  size_t len;           /* bytes (AB) or elements (TA) */
  napi_typedarray_type ta;
  struct napi_value__ *ab;
  napi_callback fn;
  prop *props;
};
struct napi_env__ { int pending; char code[64]; char msg[512]; };
struct napi_callback_info__ { size_t argc; napi_value *argv; };

static napi_value mk(vkind k) { napi_value v = calloc(1, sizeof *v); v->kind = k; return v; }
static napi_value mk_num(double d) { napi_value v = mk(V_NUM); v->num = d; return v; }
static const size_t TA_ES[] = { 1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8 };

napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data)
{
  (void)env; (void)this_arg; (void)data;
  const size_t cap = *argc;
  for (size_t i = 0; i < cap; i++) argv[i] = i < info->argc ? info->argv[i] : mk(V_UNDEF);
  *argc = info->argc;
  return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value object, size_t n, const napi_property_descriptor *p)
{
  for (size_t i = 0; i < n; i++) {
    napi_value f = mk(V_FUNC);
    f->fn = p[i].method;
    napi_set_named_property(env, object, p[i].utf8name, f);
  }
  return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value *r) { (void)env; *r = mk(V_OBJ); return napi_ok; }
napi_status napi_set_named_property(napi_env env, napi_value o, const char *name, napi_value v)
{
  (void)env;
  for (prop *p = o->props; p; p = p->next) if (!strcmp(p->name, name)) { p->v = v; return napi_ok; }
  prop *p = calloc(1, sizeof *p);
  p->name = strdup(name); p->v = v; p->next = o->props; o->props = p;
  return napi_ok;
}
napi_status napi_get_named_property(napi_env env, napi_value o, const char *name, napi_value *r)
{
  (void)env;
  for (prop *p = o->props; p; p = p->next) if (!strcmp(p->name, name)) { *r = p->v; return napi_ok; }
  *r = mk(V_UNDEF);
  return napi_ok;
}
napi_status napi_has_named_property(napi_env env, napi_value o, const char *name, bool *r)
{
  (void)env;
  *r = false;
  if (o->kind != V_OBJ) return napi_ok;
  for (prop *p = o->props; p; p = p->next) if (!strcmp(p->name, name)) *r = true;
  return napi_ok;
}
napi_status napi_create_double(napi_env env, double v, napi_value *r) { (void)env; *r = mk_num(v); return napi_ok; }
napi_status napi_create_int32(napi_env env, int32_t v, napi_value *r) { (void)env; *r = mk_num(v); return napi_ok; }
napi_status napi_create_string_utf8(napi_env env, const char *s, size_t len, napi_value *r)
{
  (void)env; (void)len;
  *r = mk(V_STR); (*r)->str = strdup(s);
  return napi_ok;
}
napi_status napi_get_value_double(napi_env env, napi_value v, double *r) { (void)env; if (v->kind != V_NUM) return napi_number_expected; *r = v->num; return napi_ok; }
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t *r) { (void)env; if (v->kind != V_NUM) return napi_number_expected; *r = (int32_t)v->num; return napi_ok; }
napi_status napi_get_value_bool(napi_env env, napi_value v, bool *r) { (void)env; if (v->kind != V_BOOL) return napi_boolean_expected; *r = v->num != 0; return napi_ok; }
napi_status napi_typeof(napi_env env, napi_value v, napi_valuetype *r)
{
  (void)env;
  switch (v->kind) {
    case V_NUM: *r = napi_number; break;
    case V_BOOL: *r = napi_boolean; break;
    case V_STR: *r = napi_string; break;
    case V_FUNC: *r = napi_function; break;
    case V_EXT: *r = napi_external; break;
    case V_UNDEF: *r = napi_undefined; break;
    default: *r = napi_object;
  }
  return napi_ok;
}
napi_status napi_get_undefined(napi_env env, napi_value *r) { (void)env; *r = mk(V_UNDEF); return napi_ok; }
napi_status napi_is_typedarray(napi_env env, napi_value v, bool *r) { (void)env; *r = v->kind == V_TA; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type *t, size_t *len, void **data, napi_value *ab, size_t *off)
{
  (void)env;
  if (v->kind != V_TA) return napi_invalid_arg;
  if (t) *t = v->ta;
  if (len) *len = v->len;
  if (data) *data = v->ptr;
  if (ab) *ab = v->ab;
  if (off) *off = 0;
  return napi_ok;
}
napi_status napi_is_arraybuffer(napi_env env, napi_value v, bool *r) { (void)env; *r = v->kind == V_AB; return napi_ok; }
napi_status napi_get_arraybuffer_info(napi_env env, napi_value v, void **data, size_t *len)
{
  (void)env;
  if (v->kind != V_AB) return napi_invalid_arg;
  *data = v->ptr; *len = v->len;
  return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t bytes, void **data, napi_value *r)
{
  (void)env;
  *r = mk(V_AB); (*r)->ptr = calloc(bytes ? bytes : 1, 1); (*r)->len = bytes;
  if (data) *data = (*r)->ptr;
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type t, size_t len, napi_value ab, size_t off, napi_value *r)
{
  (void)env;
  if (ab->kind != V_AB || off + len * TA_ES[t] > ab->len) return napi_invalid_arg;
  *r = mk(V_TA); (*r)->ta = t; (*r)->len = len; (*r)->ptr = (char *)ab->ptr + off; (*r)->ab = ab;
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void *data, napi_finalize fin, void *hint, napi_value *r)
{
  (void)env; (void)fin; (void)hint;
  *r = mk(V_EXT); (*r)->ptr = data;
  return napi_ok;
}
napi_status napi_get_value_external(napi_env env, napi_value v, void **r) { (void)env; if (v->kind != V_EXT) return napi_invalid_arg; *r = v->ptr; return napi_ok; }
napi_status napi_throw_error(napi_env env, const char *code, const char *msg)
{
  env->pending = 1;
  snprintf(env->code, sizeof env->code, "%s", code ? code : "");
  snprintf(env->msg, sizeof env->msg, "%s", msg ? msg : "");
  return napi_ok;
}

/* ---- driver ---- */
static napi_value call(napi_env env, napi_value exports, const char *name, size_t argc, napi_value *argv)
{
  napi_value f;
  napi_get_named_property(env, exports, name, &f);
  if (f->kind != V_FUNC) { fprintf(stderr, "export %s missing\n", name); exit(2); }
  struct napi_callback_info__ info = { argc, argv };
  env->pending = 0;
  return f->fn(env, &info);
}

static napi_value typed_u8(void *data, size_t n)
{
  napi_value ab = mk(V_AB); ab->ptr = data; ab->len = n;
  napi_value ta = mk(V_TA); ta->ta = napi_uint8_array; ta->len = n; ta->ptr = data; ta->ab = ab;
  return ta;
}

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: napi_host <addon.node> version|create|detect ...\n"); return 2; }
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  napi_value (*reg)(napi_env, napi_value) = (napi_value (*)(napi_env, napi_value))dlsym(h, "napi_register_module_v1");
  if (!reg) { fprintf(stderr, "no napi_register_module_v1\n"); return 2; }
  struct napi_env__ envs = { 0 };
  napi_env env = &envs;
  napi_value exports = mk(V_OBJ);
  if (reg(env, exports) != exports) { fprintf(stderr, "registration failed\n"); return 2; }
  int n_exports = 0;
  for (prop *p = exports->props; p; p = p->next) n_exports++;
  printf("exports %d\n", n_exports);

  if (!strcmp(argv[2], "version")) {
    napi_value v = call(env, exports, "version", 0, NULL);
    printf("version %s\n", v->str);
    return 0;
  }
  napi_value dev = mk_num(0);
  napi_value ctx = call(env, exports, "create", 1, &dev);
  if (env->pending) { printf("threw %s: %s\n", env->code, env->msg); return 0; }
  printf("created\n");
  if (!strcmp(argv[2], "create")) return 0;

  if (!strcmp(argv[2], "detect") && argc >= 9) {
    const int w = atoi(argv[4]), hh = atoi(argv[5]);
    unsigned char *img = malloc((size_t)w * hh);
    FILE *f = fopen(argv[3], "rb");
    if (!f || fread(img, 1, (size_t)w * hh, f) != (size_t)w * hh) { fprintf(stderr, "cannot read image\n"); return 2; }
    fclose(f);
    napi_value prm = mk(V_OBJ);
    napi_set_named_property(env, prm, "numberOfOctaves", mk_num(atoi(argv[6])));
    napi_set_named_property(env, prm, "minBlurLevel", mk_num(atof(argv[7])));
    napi_value a[6] = { ctx, typed_u8(img, (size_t)w * hh), mk_num(w), mk_num(hh), mk_num(0 /* SIFT_U8 */), prm };
    napi_value r = call(env, exports, "detect", 6, a);
    if (env->pending) { printf("threw %s: %s\n", env->code, env->msg); return 1; }
    napi_value cnt, rec, st, kp;
    napi_get_named_property(env, r, "count", &cnt);
    napi_get_named_property(env, r, "records", &rec);
    napi_get_named_property(env, r, "stats", &st);
    napi_get_named_property(env, st, "keypoints", &kp);
    printf("count %d stats.keypoints %d record_bytes %zu\n", (int)cnt->num, (int)kp->num, rec->len);
    f = fopen(argv[8], "wb");
    fwrite(rec->ptr, 80, (size_t)cnt->num, f);
    fclose(f);
    /* error convention: a bad argument must come back as a thrown error, not a crash */
    napi_value bad[6] = { ctx, typed_u8(img, 10), mk_num(w), mk_num(hh), mk_num(0), prm };
    call(env, exports, "detect", 6, bad);
    printf("short buffer %s %s\n", env->pending ? "threw" : "DID NOT THROW", env->code);
    /* stage path: buildScaleSpace -> pyramidInfo -> getLevel -> findCandidates -> refine */
    call(env, exports, "buildScaleSpace", 6, a);
    if (env->pending) { printf("threw %s: %s\n", env->code, env->msg); return 1; }
    napi_value info = call(env, exports, "pyramidInfo", 1, &ctx), oc, lv;
    napi_get_named_property(env, info, "octaves", &oc);
    napi_get_named_property(env, info, "levels", &lv);
    napi_value gl[4] = { ctx, mk_num(1), mk_num(0), mk_num(1) };
    napi_value lvl = call(env, exports, "getLevel", 4, gl), lw, lh;
    napi_get_named_property(env, lvl, "width", &lw);
    napi_get_named_property(env, lvl, "height", &lh);
    /* display product: DoG level 1 of octave 0, min/max-normalised -> Uint8ClampedArray RGBA */
    napi_value pv[6] = { ctx, mk_num(1), mk_num(0), mk_num(1), mk_num(2), mk_num(1) };
    napi_value prev = call(env, exports, "levelPreview", 6, pv), pdata;
    if (env->pending) { printf("threw %s: %s\n", env->code, env->msg); return 1; }
    napi_get_named_property(env, prev, "data", &pdata);
    {
      unsigned long long sum = 0;
      const unsigned char *pb = (const unsigned char *)pdata->ptr;
      for (size_t i = 0; i < pdata->len; i++) sum += pb[i] * (unsigned long long)(1 + (i & 3));
      printf("preview bytes %zu clamped %d checksum %llu\n", pdata->len, pdata->ta == napi_uint8_clamped_array, sum);
    }
    napi_value fb = mk(V_BOOL);
    napi_value fc[3] = { ctx, mk(V_UNDEF), fb };
    napi_value cr = call(env, exports, "findCandidates", 3, fc), ccnt, crec;
    napi_get_named_property(env, cr, "count", &ccnt);
    napi_get_named_property(env, cr, "records", &crec);
    napi_value rf[4] = { ctx, prm, crec, ccnt };
    napi_value rr = call(env, exports, "refine", 4, rf), rcnt;
    if (env->pending) { printf("threw %s: %s\n", env->code, env->msg); return 1; }
    napi_get_named_property(env, rr, "count", &rcnt);
    printf("stages octaves %d levels %d dog0_1 %dx%d candidates %d refined %d\n", (int)oc->num, (int)lv->num, (int)lw->num,
           (int)lh->num, (int)ccnt->num, (int)rcnt->num);
    return 0;
  }
  return 2;
}
