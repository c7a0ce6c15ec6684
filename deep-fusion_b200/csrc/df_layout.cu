// df_layout.cu -- weight / activation FORMAT tooling of the path (C-ABI, host memory, no device work).
//
// The reference consumes s8 weights in OIhw4i16o4i -- [O/16][I/16][kh][kw][4i][16o][4i], byte offset as in
// jit_conv_kernel.cc:333-338 and compute1x1_loop :161-175 -- and gOIhw4i16o4i (one such block per group,
// include/deepfusion.h:53-61), but ships no converter: its tests fill the blocked buffers with random bytes
// (test/test_conv_relu_pooling.cc:251-254).  Callers that hold plain oihw / goihw weights or nchw activations
// need these.  Plain loops: this runs once per model.
#include <string.h>

#include "df_common.cuh"

namespace {
inline size_t blocked_off(int o, int i, int h, int w, int ic, int kh, int kw) {
  const size_t blk = (((size_t)(o / 16) * (ic / 16) + i / 16) * kh + h) * kw + w;
  return blk * 256 + (size_t)((i % 16) / 4) * 64 + (size_t)(o % 16) * 4 + (i % 4);
}
int check_oihw(const void* a, const void* b, int groups, int oc, int ic, int kh, int kw) {
  if (!a || !b) return df::fail(DF_E_INVALID, "repack: null pointer");
  if (groups <= 0 || oc <= 0 || ic <= 0 || kh <= 0 || kw <= 0) return df::fail(DF_E_INVALID, "repack: non-positive dims");
  if (oc % 16 || ic % 16) return df::fail(DF_E_INVALID, "repack: per-group oc and ic must be multiples of 16 (got %d, %d)", oc, ic);
  return 0;
}
}  // namespace

extern "C" size_t df_wei_blocked_offset(int o, int i, int h, int w, int ic, int kh, int kw) {
  return blocked_off(o, i, h, w, ic, kh, kw);
}

// (g)oihw -> (g)OIhw4i16o4i; oc / ic are PER GROUP; groups = 1 for OIhw4i16o4i
extern "C" int df_repack_goihw_to_blocked(const int8_t* goihw, int8_t* blocked, int groups, int oc, int ic, int kh, int kw) {
  int rc = check_oihw(goihw, blocked, groups, oc, ic, kh, kw);
  if (rc) return rc;
  const size_t per_group = (size_t)oc * ic * kh * kw;
  for (int g = 0; g < groups; ++g) {
    const int8_t* src = goihw + g * per_group;
    int8_t* dst = blocked + g * per_group;
    for (int o = 0; o < oc; ++o)
      for (int i = 0; i < ic; ++i)
        for (int h = 0; h < kh; ++h)
          for (int w = 0; w < kw; ++w) dst[blocked_off(o, i, h, w, ic, kh, kw)] = src[(((size_t)o * ic + i) * kh + h) * kw + w];
  }
  return 0;
}
extern "C" int df_repack_blocked_to_goihw(const int8_t* blocked, int8_t* goihw, int groups, int oc, int ic, int kh, int kw) {
  int rc = check_oihw(blocked, goihw, groups, oc, ic, kh, kw);
  if (rc) return rc;
  const size_t per_group = (size_t)oc * ic * kh * kw;
  for (int g = 0; g < groups; ++g) {
    const int8_t* src = blocked + g * per_group;
    int8_t* dst = goihw + g * per_group;
    for (int o = 0; o < oc; ++o)
      for (int i = 0; i < ic; ++i)
        for (int h = 0; h < kh; ++h)
          for (int w = 0; w < kw; ++w) dst[(((size_t)o * ic + i) * kh + h) * kw + w] = src[blocked_off(o, i, h, w, ic, kh, kw)];
  }
  return 0;
}
extern "C" int df_repack_oihw_to_blocked(const int8_t* oihw, int8_t* blocked, int oc, int ic, int kh, int kw) {
  return df_repack_goihw_to_blocked(oihw, blocked, 1, oc, ic, kh, kw);
}
extern "C" int df_repack_blocked_to_oihw(const int8_t* blocked, int8_t* oihw, int oc, int ic, int kh, int kw) {
  return df_repack_blocked_to_goihw(blocked, oihw, 1, oc, ic, kh, kw);
}

// activations: nchw <-> nhwc for elements of `elem_bytes` (1 or 4) bytes
extern "C" int df_nchw_to_nhwc(const void* nchw, void* nhwc, int n, int c, int h, int w, int elem_bytes) {
  if (!nchw || !nhwc || n <= 0 || c <= 0 || h <= 0 || w <= 0 || (elem_bytes != 1 && elem_bytes != 4))
    return df::fail(DF_E_INVALID, "nchw_to_nhwc: bad argument");
  const char* s = static_cast<const char*>(nchw);
  char* d = static_cast<char*>(nhwc);
  for (int in = 0; in < n; ++in)
    for (int ic = 0; ic < c; ++ic)
      for (int ih = 0; ih < h; ++ih)
        for (int iw = 0; iw < w; ++iw)
          memcpy(d + ((((size_t)in * h + ih) * w + iw) * c + ic) * elem_bytes, s + ((((size_t)in * c + ic) * h + ih) * w + iw) * elem_bytes,
                 elem_bytes);
  return 0;
}
extern "C" int df_nhwc_to_nchw(const void* nhwc, void* nchw, int n, int c, int h, int w, int elem_bytes) {
  if (!nchw || !nhwc || n <= 0 || c <= 0 || h <= 0 || w <= 0 || (elem_bytes != 1 && elem_bytes != 4))
    return df::fail(DF_E_INVALID, "nhwc_to_nchw: bad argument");
  const char* s = static_cast<const char*>(nhwc);
  char* d = static_cast<char*>(nchw);
  for (int in = 0; in < n; ++in)
    for (int ic = 0; ic < c; ++ic)
      for (int ih = 0; ih < h; ++ih)
        for (int iw = 0; iw < w; ++iw)
          memcpy(d + ((((size_t)in * c + ic) * h + ih) * w + iw) * elem_bytes, s + ((((size_t)in * h + ih) * w + iw) * c + ic) * elem_bytes,
                 elem_bytes);
  return 0;
}
