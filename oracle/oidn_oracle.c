/* oidn_oracle.c -- CPU restatement of Open Image Denoise's UNet denoising path (fp32).
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oidn_b200/ may link, load or call this file; it is the
 * checker that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs compare the CUDA path against (and time as the "port" CPU baseline).
 *
 * Parity pin: the reference's CPU device cannot be built here (needs ISPC + oneTBB), so this
 * restatement is pinned against the reference's own PyTorch implementation (training/model.py,
 * training/color.py, training/tza.py, training/infer.py) run in the build container:
 * tests/golden/make_golden.py generates the vectors, tests/test_oracle_golden.py checks them.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 * Layout: tensors are HWC fp32 (the reference CPU device uses blocked Chw8c fp32; layout does not
 * change values). Weights are the TZA's fp16 values widened to fp32 (core/tensor_reorder.cpp:56-57).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* fp16 <-> fp32 (common/half.cpp:37-99: IEEE binary16, round to nearest even)                 */
/* ------------------------------------------------------------------------------------------ */
static float h2f(uint16_t h)
{
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
  if (exp == 0)
  {
    if (man == 0) bits = sign;
    else
    {
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
    }
  }
  else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4); return f;
}

static uint16_t f2h(float f)
{
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u; x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0));
  if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);
  if (x < 0x38800000u)
  {
    if (x < 0x33000000u) return (uint16_t)sign;
    int shift = 113 - (int)(x >> 23);
    uint32_t man = (x & 0x7FFFFFu) | 0x800000u;
    uint32_t r = man >> (shift + 13), rem = man & ((1u << (shift + 13)) - 1), half = 1u << (shift + 12);
    if (rem > half || (rem == half && (r & 1))) r++;
    return (uint16_t)(sign | r);
  }
  uint32_t out = (x - 0x38000000u) >> 13, rem = x & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (out & 1))) out++;
  return (uint16_t)(sign | out);
}

/* Host threads of the parallel loops. Launchers such as torchrun export OMP_NUM_THREADS=1 to their
 * workers; the timed CPU baseline (bench.py) asks for the cores it reports instead. Returns the
 * thread count in effect. */
ORO_API int oro_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

ORO_API float oro_half_to_float(uint16_t h) { return h2f(h); }
ORO_API uint16_t oro_float_to_half(float f) { return f2h(f); }

/* ------------------------------------------------------------------------------------------ */
/* TZA (core/tza.cpp:27-103; writer training/tza.py:12-108)                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
  char name[64];
  int ndims;
  uint32_t dims[4];
  char layout[8];
  char dtype; /* 'f' or 'h' */
  const void* data;
} oro_tensor;

#define ORO_MAX_TENSORS 128
typedef struct
{
  int n;
  oro_tensor t[ORO_MAX_TENSORS];
} oro_tza;

static int rd(const uint8_t** p, const uint8_t* end, void* out, size_t n)
{
  if ((size_t)(end - *p) < n) return -1;
  memcpy(out, *p, n); *p += n; return 0;
}

/* returns 0 ok; -1 corrupted; -2 unsupported version; -3 bad layout; -4 bad dtype */
ORO_API int oro_tza_parse(const void* blob, size_t size, oro_tza* out)
{
  const uint8_t* base = (const uint8_t*)blob; const uint8_t* end = base + size; const uint8_t* p = base;
  uint16_t magic; uint8_t major, minor; uint64_t table; uint32_t n;
  if (rd(&p, end, &magic, 2) || magic != 0x41D7) return -1;
  if (rd(&p, end, &major, 1) || rd(&p, end, &minor, 1)) return -1;
  if (major != 2) return -2;
  if (rd(&p, end, &table, 8) || table > size) return -1;
  p = base + table;
  if (rd(&p, end, &n, 4) || n > ORO_MAX_TENSORS) return -1;
  out->n = (int)n;
  for (uint32_t i = 0; i < n; ++i)
  {
    oro_tensor* t = &out->t[i]; memset(t, 0, sizeof(*t));
    uint16_t len; uint8_t nd; uint64_t off;
    if (rd(&p, end, &len, 2) || len >= sizeof(t->name) || rd(&p, end, t->name, len)) return -1;
    if (rd(&p, end, &nd, 1) || nd > 4) return -1;
    t->ndims = nd;
    size_t count = 1;
    for (int j = 0; j < nd; ++j) { if (rd(&p, end, &t->dims[j], 4)) return -1; count *= t->dims[j]; }
    if (rd(&p, end, t->layout, nd)) return -1;
    if (!(strcmp(t->layout, "x") == 0 || strcmp(t->layout, "oihw") == 0)) return -3;
    if (rd(&p, end, &t->dtype, 1)) return -1;
    if (t->dtype != 'f' && t->dtype != 'h') return -4;
    if (rd(&p, end, &off, 8)) return -1;
    size_t bytes = count * (t->dtype == 'f' ? 4 : 2);
    if (off > size || size - off < bytes) return -1;
    t->data = base + off;
  }
  return 0;
}

static const oro_tensor* tza_find(const oro_tza* z, const char* name)
{
  for (int i = 0; i < z->n; ++i) if (strcmp(z->t[i].name, name) == 0) return &z->t[i];
  return NULL;
}

static float tensor_get(const oro_tensor* t, size_t i)
{
  return t->dtype == 'f' ? ((const float*)t->data)[i] : h2f(((const uint16_t*)t->data)[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Transfer functions (core/color.h:10-166, core/color.cpp:9-16; twin training/color.py:49-131)  */
/* ------------------------------------------------------------------------------------------ */
enum { ORO_TF_LINEAR = 0, ORO_TF_SRGB = 1, ORO_TF_PU = 2, ORO_TF_LOG = 3 };

static float srgb_fwd(float y) { return y <= 0.0031308f ? 12.92f * y : 1.055f * powf(y, 1.f / 2.4f) + -0.055f; }
static float srgb_inv(float x) { return x <= 0.04045f ? x / 12.92f : powf((x - -0.055f) / 1.055f, 1.f / (1.f / 2.4f)); }

#define PU_A 1.41283765e+03f
#define PU_B 1.64593172e+00f
#define PU_C 4.31384981e-01f
#define PU_D -2.94139609e-03f
#define PU_E 1.92653254e-01f
#define PU_F 6.26026094e-03f
#define PU_G 9.98620152e-01f
#define PU_Y0 1.57945760e-06f
#define PU_Y1 3.22087631e-02f
#define PU_X0 2.23151711e-03f
#define PU_X1 3.70974749e-01f

static float pu_fwd(float y)
{
  if (y <= PU_Y0) return PU_A * y;
  if (y <= PU_Y1) return PU_B * powf(y, PU_C) + PU_D;
  return PU_E * logf(y + PU_F) + PU_G;
}
static float pu_inv(float x)
{
  if (x <= PU_X0) return x / PU_A;
  if (x <= PU_X1) return powf((x - PU_D) / PU_B, 1.f / PU_C);
  return expf((x - PU_G) / PU_E) - PU_F;
}

typedef struct { int type; float norm, rcp_norm; } oro_tf;

static float tf_raw_fwd(int type, float y)
{
  switch (type)
  {
  case ORO_TF_SRGB: return srgb_fwd(y);
  case ORO_TF_PU:   return pu_fwd(y);
  case ORO_TF_LOG:  return logf(y + 1.f);
  default:          return y;
  }
}

static oro_tf tf_make(int type)
{
  /* core/color.cpp:9-16: normScale = 1/forward(yMax) with normScale=1 during that call */
  oro_tf t; t.type = type; t.norm = 1.f; t.rcp_norm = 1.f;
  const float xmax = tf_raw_fwd(type, 65504.f);
  t.norm = (float)(1. / xmax); t.rcp_norm = xmax;
  return t;
}

static float tf_fwd(const oro_tf* t, float y)
{
  switch (t->type)
  {
  case ORO_TF_SRGB: return srgb_fwd(y);
  case ORO_TF_PU:   return pu_fwd(y) * t->norm;
  case ORO_TF_LOG:  return logf(y + 1.f) * t->norm;
  default:          return y;
  }
}
static float tf_inv(const oro_tf* t, float x)
{
  switch (t->type)
  {
  case ORO_TF_SRGB: return srgb_inv(x);
  case ORO_TF_PU:   return pu_inv(x * t->rcp_norm);
  case ORO_TF_LOG:  return expf(x * t->rcp_norm) - 1.f;
  default:          return x;
  }
}

ORO_API float oro_tf_forward(int type, float y) { oro_tf t = tf_make(type); return tf_fwd(&t, y); }
ORO_API float oro_tf_inverse(int type, float x) { oro_tf t = tf_make(type); return tf_inv(&t, x); }
ORO_API float oro_tf_norm_scale(int type) { oro_tf t = tf_make(type); return t.norm; }

/* ------------------------------------------------------------------------------------------ */
/* Images (core/image.h:14-120, core/image_accessor.h:16-94)                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
  void* ptr;          /* NULL = image not set */
  int is_half;        /* 0 fp32, 1 fp16 */
  int C, H, W;        /* 1..3 channels */
  size_t pixel_stride, row_stride; /* bytes */
} oro_image;

static void img_get3(const oro_image* im, int h, int w, float v[3])
{
  const uint8_t* px = (const uint8_t*)im->ptr + (size_t)h * im->row_stride + (size_t)w * im->pixel_stride;
  float c[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < im->C; ++i) c[i] = im->is_half ? h2f(((const uint16_t*)px)[i]) : ((const float*)px)[i];
  /* image_accessor.h:36-41: C==2 -> (x,y,y), C==1 -> (x,x,x) */
  v[0] = c[0]; v[1] = im->C >= 2 ? c[1] : c[0]; v[2] = im->C == 3 ? c[2] : (im->C == 2 ? c[1] : c[0]);
}

static void img_set3(const oro_image* im, int h, int w, const float v[3])
{
  uint8_t* px = (uint8_t*)im->ptr + (size_t)h * im->row_stride + (size_t)w * im->pixel_stride;
  for (int i = 0; i < im->C; ++i)
  {
    if (im->is_half) ((uint16_t*)px)[i] = f2h(v[i]);
    else ((float*)px)[i] = v[i];
  }
}

static float nan_to_zero(float x) { return isnan(x) ? 0.f : x; }
static float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* ------------------------------------------------------------------------------------------ */
/* Autoexposure (core/autoexposure.h:14-19, devices/cpu/cpu_autoexposure.cpp:22-64,            */
/* cpu_autoexposure.ispc:8-25; twin training/color.py:138-173)                                  */
/* ------------------------------------------------------------------------------------------ */
ORO_API float oro_autoexposure(const oro_image* color)
{
  const int H = color->H, W = color->W;
  const int nbh = (H + 15) / 16, nbw = (W + 15) / 16;
  double sum = 0.; long count = 0;
  for (int i = 0; i < nbh; ++i)
    for (int j = 0; j < nbw; ++j)
    {
      const int bh = (int)((long)i * H / nbh), eh = (int)((long)(i + 1) * H / nbh);
      const int bw = (int)((long)j * W / nbw), ew = (int)((long)(j + 1) * W / nbw);
      float L = 0.f;
      for (int h = bh; h < eh; ++h)
        for (int w = bw; w < ew; ++w)
        {
          float c[3]; img_get3(color, h, w, c);
          for (int k = 0; k < 3; ++k) c[k] = clampf(nan_to_zero(c[k]), 0.f, FLT_MAX);
          L += 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2];
        }
      L /= (float)((eh - bh) * (ew - bw));
      if (L > 1e-8f) { sum += log2f(L); count++; }
    }
  return count > 0 ? 0.18f / exp2f((float)(sum / (double)count)) : 1.f;
}

/* ------------------------------------------------------------------------------------------ */
/* Input / output process (devices/cpu/cpu_input_process.isph:31-136,                           */
/* cpu_output_process.isph:29-70, core/tile.h)                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int hSrcBegin, wSrcBegin, hDstBegin, wDstBegin, H, W; } oro_tile;

/* dst: HWC fp32 [TH][TW][C], C = 3*(#images); zero outside the tile footprint */
ORO_API void oro_input_process(const oro_image* color, const oro_image* albedo, const oro_image* normal,
                               const oro_tile* tile, int tf_type, int hdr, int snorm, float input_scale,
                               float* dst, int TH, int TW, int C)
{
  const oro_tf tf = tf_make(tf_type);
  /* the "input" image is color, else albedo, else normal (core/unet_filter.cpp:368, input_process.cpp) */
  const oro_image* input = (color && color->ptr) ? color : ((albedo && albedo->ptr) ? albedo : normal);
  const int has_alb = color && color->ptr && albedo && albedo->ptr;
  const int has_nrm = has_alb && normal && normal->ptr;
  memset(dst, 0, (size_t)TH * TW * C * sizeof(float));
  #pragma omp parallel for schedule(static)
  for (int hd = 0; hd < TH; ++hd)
  {
    const int h = hd - tile->hDstBegin;
    if (h < 0 || h >= tile->H) continue;
    for (int w = 0; w < tile->W; ++w)
    {
      const int hs = h + tile->hSrcBegin, ws = w + tile->wSrcBegin, wd = w + tile->wDstBegin;
      float* o = dst + ((size_t)hd * TW + wd) * C;
      float v[3]; img_get3(input, hs, ws, v);
      for (int k = 0; k < 3; ++k)
      {
        float x = v[k] * input_scale;
        x = clampf(nan_to_zero(x), snorm ? -1.f : 0.f, hdr ? FLT_MAX : 1.f);
        if (snorm) x = x * 0.5f + 0.5f;
        o[k] = tf_fwd(&tf, x);
      }
      if (has_alb)
      {
        img_get3(albedo, hs, ws, v);
        for (int k = 0; k < 3; ++k) o[3 + k] = clampf(nan_to_zero(v[k]), 0.f, 1.f);
        if (has_nrm)
        {
          img_get3(normal, hs, ws, v);
          for (int k = 0; k < 3; ++k) o[6 + k] = clampf(nan_to_zero(v[k]), -1.f, 1.f) * 0.5f + 0.5f;
        }
      }
    }
  }
}

/* src: HWC fp32 [TH][TW][C>=3] */
ORO_API void oro_output_process(const float* src, int TH, int TW, int C, const oro_tile* tile, int tf_type,
                                int hdr, int snorm, float input_scale, const oro_image* out)
{
  (void)TH;
  const oro_tf tf = tf_make(tf_type);
  const float output_scale = input_scale != 0.f ? 1.f / input_scale : 0.f; /* core/color.h:95-123 */
  #pragma omp parallel for schedule(static)
  for (int h = 0; h < tile->H; ++h)
    for (int w = 0; w < tile->W; ++w)
    {
      const float* s = src + ((size_t)(h + tile->hSrcBegin) * TW + (w + tile->wSrcBegin)) * C;
      float v[3];
      for (int k = 0; k < 3; ++k) v[k] = tf_inv(&tf, clampf(nan_to_zero(s[k]), 0.f, FLT_MAX));
      if (out->C == 1) { const float m = (v[0] + v[1] + v[2]) * (1.f / 3.f); v[0] = v[1] = v[2] = m; }
      for (int k = 0; k < 3; ++k)
      {
        if (snorm) { v[k] = v[k] * 2.f - 1.f; v[k] = fmaxf(v[k], -1.f); }
        if (!hdr) v[k] = fminf(v[k], 1.f);
        v[k] *= output_scale;
      }
      img_set3(out, h + tile->hDstBegin, w + tile->wDstBegin, v);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Conv / Pool / Upsample (core/conv.cpp, devices/cpu/cpu_conv.ispc:34-127,                     */
/* cpu_pool.isph:15-34, cpu_upsample.isph:15-33)                                                */
/* ------------------------------------------------------------------------------------------ */
/* src HWC [H][W][Cin], weights OIHW (fp32, widened), dst HWC [H][W][Cout]; 3x3, pad 1, fp32 FMA
 * accumulation starting from the bias, order kh -> kw -> ci as in cpu_conv.ispc. */
ORO_API void oro_conv3x3(const float* src, int H, int W, int Cin, const float* w_oihw, const float* bias,
                         int Cout, int relu, float* dst)
{
  /* repack weights to [kh][kw][ci][co] so the inner loop runs over contiguous output channels */
  float* wt = (float*)malloc((size_t)9 * Cin * Cout * sizeof(float));
  for (int o = 0; o < Cout; ++o)
    for (int i = 0; i < Cin; ++i)
      for (int k = 0; k < 9; ++k)
        wt[((size_t)k * Cin + i) * Cout + o] = w_oihw[((size_t)o * Cin + i) * 9 + k];
  /* zero-padded copy of the source removes the border tests */
  const int PW = W + 2;
  float* pad = (float*)calloc((size_t)(H + 2) * PW * Cin, sizeof(float));
  for (int y = 0; y < H; ++y)
    memcpy(pad + ((size_t)(y + 1) * PW + 1) * Cin, src + (size_t)y * W * Cin, (size_t)W * Cin * sizeof(float));

  #pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
  {
    float acc[4][256];
    for (int x0 = 0; x0 < W; x0 += 4)
    {
      const int nx = W - x0 < 4 ? W - x0 : 4;
      for (int p = 0; p < 4; ++p) for (int o = 0; o < Cout; ++o) acc[p][o] = bias[o];
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
        {
          const float* wk = wt + (size_t)(kh * 3 + kw) * Cin * Cout;
          const float* s0 = pad + ((size_t)(y + kh) * PW + (x0 + kw)) * Cin;
          if (nx == 4)
          {
            for (int i = 0; i < Cin; ++i)
            {
              const float a0 = s0[i], a1 = s0[Cin + i], a2 = s0[2 * Cin + i], a3 = s0[3 * Cin + i];
              const float* wr = wk + (size_t)i * Cout;
              for (int o = 0; o < Cout; ++o)
              {
                const float wv = wr[o];
                acc[0][o] += a0 * wv; acc[1][o] += a1 * wv; acc[2][o] += a2 * wv; acc[3][o] += a3 * wv;
              }
            }
          }
          else
          {
            for (int p = 0; p < nx; ++p)
              for (int i = 0; i < Cin; ++i)
              {
                const float a = s0[(size_t)p * Cin + i];
                const float* wr = wk + (size_t)i * Cout;
                for (int o = 0; o < Cout; ++o) acc[p][o] += a * wr[o];
              }
          }
        }
      for (int p = 0; p < nx; ++p)
      {
        float* d = dst + ((size_t)y * W + x0 + p) * Cout;
        for (int o = 0; o < Cout; ++o) d[o] = relu ? fmaxf(acc[p][o], 0.f) : acc[p][o];
      }
    }
  }
  free(pad); free(wt);
}

ORO_API void oro_pool2x2(const float* src, int H, int W, int C, float* dst)
{
  const int Ho = H / 2, Wo = W / 2;
  #pragma omp parallel for schedule(static)
  for (int y = 0; y < Ho; ++y)
    for (int x = 0; x < Wo; ++x)
      for (int c = 0; c < C; ++c)
      {
        const float a = src[((size_t)(2 * y) * W + 2 * x) * C + c], b = src[((size_t)(2 * y) * W + 2 * x + 1) * C + c];
        const float d = src[((size_t)(2 * y + 1) * W + 2 * x) * C + c], e = src[((size_t)(2 * y + 1) * W + 2 * x + 1) * C + c];
        dst[((size_t)y * Wo + x) * C + c] = fmaxf(fmaxf(a, b), fmaxf(d, e));
      }
}

ORO_API void oro_upsample2x(const float* src, int H, int W, int C, float* dst)
{
  #pragma omp parallel for schedule(static)
  for (int y = 0; y < 2 * H; ++y)
    for (int x = 0; x < 2 * W; ++x)
      memcpy(dst + ((size_t)y * 2 * W + x) * C, src + ((size_t)(y / 2) * W + x / 2) * C, (size_t)C * sizeof(float));
}

static float* concat_hwc(const float* a, int Ca, const float* b, int Cb, int H, int W)
{
  float* o = (float*)malloc((size_t)H * W * (Ca + Cb) * sizeof(float));
  #pragma omp parallel for schedule(static)
  for (long p = 0; p < (long)H * W; ++p)
  {
    memcpy(o + p * (Ca + Cb), a + p * Ca, (size_t)Ca * sizeof(float));
    memcpy(o + p * (Ca + Cb) + Ca, b + p * Cb, (size_t)Cb * sizeof(float));
  }
  return o;
}

/* ------------------------------------------------------------------------------------------ */
/* UNet graphs (core/unet_filter.cpp:468-531; twin training/model.py:55-260)                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float* d; int C, H, W; } tens;

static tens conv_layer(const oro_tza* z, const char* name, tens in, int relu, int* err)
{
  char wn[80], bn[80];
  snprintf(wn, sizeof wn, "%s.weight", name); snprintf(bn, sizeof bn, "%s.bias", name);
  const oro_tensor* wt = tza_find(z, wn); const oro_tensor* bt = tza_find(z, bn);
  tens out = {NULL, 0, in.H, in.W};
  if (!wt || !bt || wt->ndims != 4 || (int)wt->dims[1] != in.C || wt->dims[2] != 3 || wt->dims[3] != 3 ||
      bt->ndims != 1 || bt->dims[0] != wt->dims[0])
  { *err = 1; return out; }
  const int O = (int)wt->dims[0];
  const size_t nw = (size_t)O * in.C * 9;
  float* w = (float*)malloc(nw * sizeof(float)); float* b = (float*)malloc((size_t)O * sizeof(float));
  for (size_t i = 0; i < nw; ++i) w[i] = tensor_get(wt, i);
  for (int i = 0; i < O; ++i) b[i] = tensor_get(bt, (size_t)i);
  out.C = O; out.d = (float*)malloc((size_t)in.H * in.W * O * sizeof(float));
  oro_conv3x3(in.d, in.H, in.W, in.C, w, b, O, relu, out.d);
  free(w); free(b);
  return out;
}

static tens pool_layer(tens in)
{
  tens o = {(float*)malloc((size_t)(in.H / 2) * (in.W / 2) * in.C * sizeof(float)), in.C, in.H / 2, in.W / 2};
  oro_pool2x2(in.d, in.H, in.W, in.C, o.d); return o;
}
static tens up_layer(tens in)
{
  tens o = {(float*)malloc((size_t)(in.H * 2) * (in.W * 2) * in.C * sizeof(float)), in.C, in.H * 2, in.W * 2};
  oro_upsample2x(in.d, in.H, in.W, in.C, o.d); return o;
}
static tens cat_layer(tens a, tens b)
{
  tens o = {concat_hwc(a.d, a.C, b.d, b.C, a.H, a.W), a.C + b.C, a.H, a.W}; return o;
}
#define FREE(t) do { free((t).d); (t).d = NULL; } while (0)

/* in: HWC fp32 [H][W][C], H,W multiples of 16. Returns 3-channel HWC output (malloc'd) or NULL. */
static float* unet_forward(const oro_tza* z, const float* in, int H, int W, int C, int* outC)
{
  int err = 0;
  const int large = tza_find(z, "enc_conv1b.weight") != NULL; /* core/unet_filter.cpp:263 */
  tens input = {(float*)in, C, H, W};
  tens x, t, pool1, pool2, pool3, u, c;
  if (!large)
  {
    x = conv_layer(z, "enc_conv0", input, 1, &err); if (err) return NULL;
    t = conv_layer(z, "enc_conv1", x, 1, &err); FREE(x); pool1 = pool_layer(t); FREE(t);
    t = conv_layer(z, "enc_conv2", pool1, 1, &err); pool2 = pool_layer(t); FREE(t);
    t = conv_layer(z, "enc_conv3", pool2, 1, &err); pool3 = pool_layer(t); FREE(t);
    t = conv_layer(z, "enc_conv4", pool3, 1, &err); x = pool_layer(t); FREE(t);
  }
  else
  {
    x = conv_layer(z, "enc_conv1a", input, 1, &err); if (err) return NULL;
    t = conv_layer(z, "enc_conv1b", x, 1, &err); FREE(x); pool1 = pool_layer(t); FREE(t);
    x = conv_layer(z, "enc_conv2a", pool1, 1, &err);
    t = conv_layer(z, "enc_conv2b", x, 1, &err); FREE(x); pool2 = pool_layer(t); FREE(t);
    x = conv_layer(z, "enc_conv3a", pool2, 1, &err);
    t = conv_layer(z, "enc_conv3b", x, 1, &err); FREE(x); pool3 = pool_layer(t); FREE(t);
    x = conv_layer(z, "enc_conv4a", pool3, 1, &err);
    t = conv_layer(z, "enc_conv4b", x, 1, &err); FREE(x); x = pool_layer(t); FREE(t);
  }
  if (err) return NULL;
  t = conv_layer(z, "enc_conv5a", x, 1, &err); FREE(x);
  x = conv_layer(z, "enc_conv5b", t, 1, &err); FREE(t);
  if (err) return NULL;
  const char* names[4][2] = {{"dec_conv4a", "dec_conv4b"}, {"dec_conv3a", "dec_conv3b"},
                             {"dec_conv2a", "dec_conv2b"}, {"dec_conv1a", "dec_conv1b"}};
  tens skips[4] = {pool3, pool2, pool1, input};
  for (int l = 0; l < 4; ++l)
  {
    u = up_layer(x); FREE(x);
    c = cat_layer(u, skips[l]); FREE(u);   /* order: (upsampled decoder, skip) model.py:134-149 */
    if (l < 3) FREE(skips[l]);
    t = conv_layer(z, names[l][0], c, 1, &err); FREE(c);
    if (err) return NULL;
    x = conv_layer(z, names[l][1], t, 1, &err); FREE(t);
    if (err) return NULL;
  }
  /* last conv has ReLU in the C++ graph (unet_filter.cpp:495,528) */
  t = conv_layer(z, large ? "dec_conv1c" : "dec_conv0", x, 1, &err); FREE(x);
  if (err) return NULL;
  *outC = t.C;
  return t.d;
}

/* Raw network forward on an HWC fp32 tensor (unit-test entry). out must hold H*W*outC floats. */
ORO_API int oro_unet_forward(const void* tza_blob, size_t tza_size, const float* in, int H, int W, int C,
                             float* out, int out_capacity_floats)
{
  oro_tza z; int rc = oro_tza_parse(tza_blob, tza_size, &z); if (rc) return rc;
  int oc = 0; float* o = unet_forward(&z, in, H, W, C, &oc);
  if (!o) return -10;
  if ((size_t)H * W * oc > (size_t)out_capacity_floats) { free(o); return -11; }
  memcpy(out, o, (size_t)H * W * oc * sizeof(float)); free(o);
  return oc;
}

/* ------------------------------------------------------------------------------------------ */
/* Tile planner + per-tile execute loop (core/unet_filter.cpp:254-335 init, :198-241 execute)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
  int H, W, tileH, tileW, tilePadH, tilePadW, tileCountH, tileCountW, tileOverlap, tileAlignment;
} oro_tiling;

static int round_up_i(int a, int b) { return (a + b - 1) / b * b; }
static int round_up3(int a, int b, int c) { return round_up_i(a - c, b) + c; } /* common/platform.h:203-207 style */
static int ceil_div_i(int a, int b) { return (a + b - 1) / b; }
static int clamp_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int max_i(int a, int b) { return a > b ? a : b; }
static int min_i(int a, int b) { return a < b ? a : b; }

/* max_tile_pixels: the reference uses 2160*2160 when maxMemoryMB < 0 (unet_filter.h:39, .cpp:300).
 * The memory-size test of buildModel() is device dependent and not restated; pass a smaller
 * max_tile_pixels to force more tiles (what maxMemoryMB=0 does in oidnTest.cpp:717-775). */
ORO_API void oro_plan_tiles(int H, int W, int large, int dev_alignment, int num_subdevices, long max_tile_pixels,
                            oro_tiling* t)
{
  const int minTileAlignment = 16;
  const int rf = large ? 202 : 174;
  int a = minTileAlignment, b = dev_alignment > 0 ? dev_alignment : 1, g = a, h = b;
  while (h) { int r = g % h; g = h; h = r; }
  t->tileAlignment = a / g * b;                       /* lcm */
  t->tileOverlap = round_up_i(rf / 2, t->tileAlignment);
  t->H = H; t->W = W;
  t->tileH = round_up_i(H, minTileAlignment);
  t->tileW = round_up_i(W, minTileAlignment);
  t->tilePadH = t->tileH % t->tileAlignment;
  t->tilePadW = t->tileW % t->tileAlignment;
  t->tileCountH = t->tileCountW = 1;
  const int minTileDim = max_i(4 * t->tileOverlap, 768);
  const int minTileH = round_up3(minTileDim, t->tileAlignment, t->tilePadH);
  const int minTileW = round_up3(minTileDim, t->tileAlignment, t->tilePadW);
  const int ov2H = 2 * t->tileOverlap + t->tilePadH, ov2W = 2 * t->tileOverlap + t->tilePadW;
  while ((t->tileCountH * t->tileCountW) % num_subdevices != 0 || (long)t->tileH * t->tileW > max_tile_pixels)
  {
    if (t->tileH > minTileH && t->tileH > t->tileW)
    {
      const int newH = ceil_div_i(H + ov2H * t->tileCountH, t->tileCountH + 1);
      t->tileH = clamp_i(round_up3(newH, t->tileAlignment, t->tilePadH), minTileH, t->tileH - t->tileAlignment);
      t->tileCountH = max_i(ceil_div_i(H - ov2H, t->tileH - ov2H), 1);
    }
    else if (t->tileW > minTileW)
    {
      const int newW = ceil_div_i(W + ov2W * t->tileCountW, t->tileCountW + 1);
      t->tileW = clamp_i(round_up3(newW, t->tileAlignment, t->tilePadW), minTileW, t->tileW - t->tileAlignment);
      t->tileCountW = max_i(ceil_div_i(W - ov2W, t->tileW - ov2W), 1);
    }
    else break; /* cannot divide further */
  }
}

typedef struct
{
  int filter;        /* 0 RT, 1 RTLightmap */
  int hdr, srgb, directional;
  float input_scale; /* NaN = auto (autoexposure when hdr, else 1) */
  long max_tile_pixels; /* <=0: reference default 2160*2160 */
  int num_subdevices;   /* >=1 */
} oro_params;

typedef struct { int tileH, tileW, tileCountH, tileCountW, tileOverlap, large; float input_scale; } oro_stats;

/* The whole filter: RT / RTLightmap execute on the CPU. Returns 0 on success. */
ORO_API int oro_filter_execute(const void* tza_blob, size_t tza_size, const oro_image* color,
                               const oro_image* albedo, const oro_image* normal, const oro_image* output,
                               const oro_params* prm, oro_stats* stats)
{
  oro_tza z; int rc = oro_tza_parse(tza_blob, tza_size, &z); if (rc) return rc;
  const int has_color = color && color->ptr, has_alb = albedo && albedo->ptr, has_nrm = normal && normal->ptr;
  if (!has_color && !has_alb && !has_nrm) return -20;
  if (!output || !output->ptr) return -21;
  const int H = output->H, W = output->W;
  if (H <= 0 || W <= 0) return 0;
  int hdr = prm->hdr, srgb = prm->srgb, directional = prm->directional;
  if (prm->filter == 1) hdr = !directional; /* core/rtlightmap_filter.cpp:58-62, ctor hdr=true */
  if (directional && (hdr || srgb)) return -22;
  if (hdr && srgb) return -23;
  /* transfer function: core/rt_filter.cpp:63-71, core/rtlightmap_filter.cpp:24-30 */
  int tf_type;
  if (prm->filter == 1) tf_type = hdr ? ORO_TF_LOG : ORO_TF_LINEAR;
  else tf_type = (srgb || (!has_color && has_nrm)) ? ORO_TF_LINEAR : (hdr ? ORO_TF_PU : ORO_TF_SRGB);
  const int snorm = directional || (!has_color && has_nrm); /* unet_filter.cpp:551 */
  int inputC = 0;
  if (has_color) inputC += 3;
  if (has_alb) inputC += 3;
  if (has_nrm) inputC += 3;
  const int large = tza_find(&z, "enc_conv1b.weight") != NULL;

  float scale = prm->input_scale;
  if (isnan(scale)) scale = hdr ? oro_autoexposure(color) : 1.f; /* unet_filter.cpp:172-189 */

  oro_tiling t;
  oro_plan_tiles(H, W, large, 1, prm->num_subdevices > 0 ? prm->num_subdevices : 1,
                 prm->max_tile_pixels > 0 ? prm->max_tile_pixels : 2160L * 2160L, &t);
  if (stats)
  {
    stats->tileH = t.tileH; stats->tileW = t.tileW; stats->tileCountH = t.tileCountH;
    stats->tileCountW = t.tileCountW; stats->tileOverlap = t.tileOverlap; stats->large = large;
    stats->input_scale = scale;
  }

  /* in-place + tiled needs a temporary output (unet_filter.cpp:583-590); the oracle always
   * renders into a temporary and copies at the end, which is equivalent. */
  const size_t esz = output->is_half ? 2 : 4;
  oro_image tmp = *output;
  tmp.pixel_stride = (size_t)output->C * esz; tmp.row_stride = tmp.pixel_stride * W;
  tmp.ptr = malloc(tmp.row_stride * H);

  float* in = (float*)malloc((size_t)t.tileH * t.tileW * inputC * sizeof(float));
  const oro_image* c0 = has_color ? color : NULL;
  for (int i = 0; i < t.tileCountH; ++i)
  {
    const int ovH = 2 * t.tileOverlap + t.tilePadH;
    const int h = i * (t.tileH - ovH);
    const int obH = i > 0 ? t.tileOverlap : 0, oeH = i < t.tileCountH - 1 ? t.tileOverlap + t.tilePadH : 0;
    const int tileH1 = min_i(H - h, t.tileH), tileH2 = tileH1 - obH - oeH;
    const int alignH = t.tileH - round_up_i(tileH1, 16);
    for (int j = 0; j < t.tileCountW; ++j)
    {
      const int ovW = 2 * t.tileOverlap + t.tilePadW;
      const int w = j * (t.tileW - ovW);
      const int obW = j > 0 ? t.tileOverlap : 0, oeW = j < t.tileCountW - 1 ? t.tileOverlap + t.tilePadW : 0;
      const int tileW1 = min_i(W - w, t.tileW), tileW2 = tileW1 - obW - oeW;
      const int alignW = t.tileW - round_up_i(tileW1, 16);
      const oro_tile it = {h, w, alignH, alignW, tileH1, tileW1};
      const oro_tile ot = {alignH + obH, alignW + obW, h + obH, w + obW, tileH2, tileW2};
      oro_input_process(c0, has_color ? albedo : (has_alb ? albedo : NULL), has_color ? normal : (has_nrm ? normal : NULL),
                        &it, tf_type, hdr, snorm, scale, in, t.tileH, t.tileW, inputC);
      int oc = 0;
      float* o = unet_forward(&z, in, t.tileH, t.tileW, inputC, &oc);
      if (!o) { free(in); free(tmp.ptr); return -10; }
      oro_output_process(o, t.tileH, t.tileW, oc, &ot, tf_type, hdr, snorm, scale, &tmp);
      free(o);
    }
  }
  free(in);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      memcpy((uint8_t*)output->ptr + (size_t)y * output->row_stride + (size_t)x * output->pixel_stride,
             (uint8_t*)tmp.ptr + (size_t)y * tmp.row_stride + (size_t)x * tmp.pixel_stride, (size_t)output->C * esz);
  free(tmp.ptr);
  return 0;
}
