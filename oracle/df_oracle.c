/* df_oracle.c -- scalar CPU oracle.  TEST INFRASTRUCTURE ONLY (see df_oracle.h).
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * Every f32 operation below is a single IEEE-754 binary32 operation in round-to-nearest-even,
 * matching one x86 vector instruction of the reference's JIT output.
 */
#include "df_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <float.h>
#include <string.h>

/* ------------------------------------------------------------------ element semantics */

/* vcvtps2dq zmm|T_r{n,d}_sae (jit_conv_kernel.cc:105-112, :267-274): NaN and anything that
 * does not fit s32 after rounding become the "integer indefinite" 0x80000000. */
int32_t dfo_cvt_f32_s32(float t, int round_mode) {
  if (t != t) return INT32_MIN;
  float r = (round_mode == DFO_DOWN) ? floorf(t) : rintf(t); /* rintf: RN-even by default */
  if (!(r < 2147483648.0f) || r < -2147483648.0f) return INT32_MIN;
  return (int32_t)r;
}

/* vmaxps(dst, zero, t) (jit_conv_kernel.cc:103, :265): x86 MAXPS returns the SECOND source
 * when either input is NaN or both are zero, otherwise the larger. */
float dfo_relu_f32(float t) { return (0.0f > t) ? 0.0f : t; }

/* vpmovusdb: the dword is treated as UNSIGNED and saturated to 255 */
uint8_t dfo_usat8(int32_t v) { return ((uint32_t)v > 255u) ? 255 : (uint8_t)v; }
/* vpmovsdb: signed saturation */
int8_t dfo_ssat8(int32_t v) { return v > 127 ? 127 : (v < -128 ? -128 : (int8_t)v); }

static float load_bias_f32(int dt, const void *bia, int idx) {
  switch (dt) { /* jit_conv_kernel.cc:69-85, :238-254 */
    case DFO_F32: return ((const float *)bia)[idx];
    case DFO_S32: return (float)((const int32_t *)bia)[idx]; /* vcvtdq2ps, RN */
    case DFO_S8: return (float)((const int8_t *)bia)[idx];   /* vpmovsxbd + vcvtdq2ps */
    case DFO_U8: return (float)((const uint8_t *)bia)[idx];  /* vpmovzxbd + vcvtdq2ps */
    default: return 0.0f;
  }
}

/* vcvtdq2ps ; [vaddps bias] ; vmulps scale  -- three separately rounded operations
 * (jit_conv_kernel.cc:96-100, :259-263).  volatile keeps the compiler from contracting. */
float dfo_epilogue_f32(int32_t acc, int bia_dt, const void *bia, int idx, float scale) {
  volatile float t = (float)acc;
  if (bia_dt != DFO_UNDEF && bia) {
    volatile float b = load_bias_f32(bia_dt, bia, idx);
    t = t + b;
  }
  t = t * scale;
  return t;
}

/* ----------------------------------------------------------------------------- helpers */

size_t dfo_wei_off(int o, int i, int h, int w, int ic, int kh, int kw) {
  /* [O/16][I/16][kh][kw][4i][16o][4i]  (kernel_offset, jit_conv_kernel.cc:333-338 and the
   * kh step :326/:384; 1x1 weights are the kh=kw=1 case, :161-175) */
  size_t nb_ic = (size_t)ic / 16;
  size_t ob = (size_t)o / 16, ib = (size_t)i / 16;
  size_t blk = ((ob * nb_ic + ib) * kh + h) * kw + w;
  return blk * 256 + (size_t)((i % 16) / 4) * 64 + (size_t)(o % 16) * 4 + (size_t)(i % 4);
}

void dfo_repack_oihw(const int8_t *oihw, int8_t *blocked, int oc, int ic, int kh, int kw) {
  for (int o = 0; o < oc; ++o)
    for (int i = 0; i < ic; ++i)
      for (int h = 0; h < kh; ++h)
        for (int w = 0; w < kw; ++w)
          blocked[dfo_wei_off(o, i, h, w, ic, kh, kw)] =
              oihw[(((size_t)o * ic + i) * kh + h) * kw + w];
}

int dfo_conv_output_size(int image, int kernel, int stride, int padding) {
  return (image + 2 * padding - kernel) / stride + 1; /* util/math_func.cc:22-24 */
}

int dfo_dividable_of(int val, const int *divisors, int n) {
  for (int k = 0; k < n; ++k)
    if (val % divisors[k] == 0) return divisors[k];
  return 1; /* deepfusion_utils.h:116-132 */
}

int dfo_find_dividable(int val, int divisor) {
  if (divisor <= 1) return 1; /* deepfusion_utils.h:134-148 */
  if (divisor > val) return val;
  while (divisor > 1 && val % divisor != 0) --divisor;
  return divisor;
}

void dfo_balance211(long n, int team, int tid, long *start, long *end) {
  /* deepfusion_utils.h:190-209 */
  if (team <= 1 || n == 0) {
    *start = 0;
    *end = n;
    return;
  }
  long n1 = (n + team - 1) / team, n2 = n1 - 1, T1 = n - n2 * team;
  long my = tid < T1 ? n1 : n2;
  *start = tid <= T1 ? tid * n1 : T1 * n1 + (tid - T1) * n2;
  *end = *start + my;
}

void dfo_conv_blocking(int ic, int oc, int ow, int kh, int kw, int out[4]) {
  /* jit_conv_kernel.cc:643-655; ker_reg_base_idx = 28 (jit_conv_kernel.h:54-56) */
  int nb_ic = ic / 16, nb_oc = oc / 16;
  const int d8[] = {8, 4, 2, 1}, d4[] = {4, 2, 1};
  int nb_ic_blocking = dfo_dividable_of(nb_ic, d8, 4);
  if (kh >= 7 || kw >= 7) nb_ic_blocking = dfo_dividable_of(nb_ic, d4, 3);
  int nb_oc_blocking = nb_oc > 4 ? 4 : nb_oc;
  if (nb_oc % nb_oc_blocking != 0) nb_oc_blocking = dfo_find_dividable(nb_oc, nb_oc_blocking);
  int ur_w = 28 / (nb_oc_blocking + 1);
  if (ow < ur_w) ur_w = ow;
  out[0] = nb_ic_blocking;
  out[1] = nb_oc_blocking;
  out[2] = ur_w;
  out[3] = ow % ur_w;
}

static int is_io_dt(int dt) { return dt == DFO_F32 || dt == DFO_S32 || dt == DFO_S8 || dt == DFO_U8; }

int dfo_conv_check(const dfo_conv_desc *d) {
  if (!d) return -1;
  if (d->n <= 0 || d->ih <= 0 || d->iw <= 0 || d->kh <= 0 || d->kw <= 0 || d->sh <= 0 ||
      d->sw <= 0 || d->ph < 0 || d->pw < 0)
    return -2;
  if (!is_io_dt(d->dst_dt)) return -3;                               /* :534-538 */
  if (d->bia0_dt != DFO_UNDEF && !is_io_dt(d->bia0_dt)) return -4;   /* :539-543 */
  if (d->bia1_dt != DFO_UNDEF && !is_io_dt(d->bia1_dt)) return -4;
  if (d->ic <= 0 || d->oc <= 0 || d->ic % 16 || d->oc % 16) return -5; /* :590 */
  if (d->oc1 < 0 || d->oc1 % 16) return -6;                          /* :616 */
  if (d->round0 != DFO_NEAREST && d->round0 != DFO_DOWN) return -7;
  if (d->round1 != DFO_NEAREST && d->round1 != DFO_DOWN) return -7;
  if (d->nscale0 != 1 && d->nscale0 != d->oc) return -8;             /* :665 */
  if (d->oc1 && d->nscale1 != 1 && d->nscale1 != d->oc1) return -8;  /* :668 */
  int oh = dfo_conv_output_size(d->ih, d->kh, d->sh, d->ph);
  int ow = dfo_conv_output_size(d->iw, d->kw, d->sw, d->pw);
  if (oh <= 0 || ow <= 0) return -9;
  int blk[4];
  dfo_conv_blocking(d->ic, d->oc, ow, d->kh, d->kw, blk);
  int ur_w = blk[2], tail = blk[3];
  int r_pad_no_tail = (ow - tail - 1) * d->sw + d->kw - d->iw - d->pw; /* :657-661 */
  if (r_pad_no_tail < 0) r_pad_no_tail = 0;
  if (d->pw > ur_w || r_pad_no_tail > ur_w) return -10;
  return 0;
}

/* ------------------------------------------------------------------------------- conv */

/* `w0c` = the blocked weights gathered once to [o][kh][kw][i] so the inner loop is contiguous */
static int32_t conv0_acc(const dfo_conv_desc *d, const uint8_t *src, const int8_t *w0c, int n,
                         int oh, int ow, int o) {
  /* C1: exact s32 sum of u8*s8 over the kh x kw x ic window; taps outside the image are
   * simply not executed (kh_padding / get_ow_start / get_ow_end, op_conv.cc:218-220,
   * jit_conv_kernel.h:120-127), i.e. contribute zero. */
  int32_t acc = 0;
  for (int kh = 0; kh < d->kh; ++kh) {
    int ih = oh * d->sh - d->ph + kh;
    if (ih < 0 || ih >= d->ih) continue;
    for (int kw = 0; kw < d->kw; ++kw) {
      int iw = ow * d->sw - d->pw + kw;
      if (iw < 0 || iw >= d->iw) continue;
      const uint8_t *px = src + (((size_t)n * d->ih + ih) * d->iw + iw) * d->ic;
      const int8_t *wr = w0c + (((size_t)o * d->kh + kh) * d->kw + kw) * d->ic;
      for (int i = 0; i < d->ic; ++i) acc += (int32_t)px[i] * (int32_t)wr[i];
    }
  }
  return acc;
}

static void store_dst(int dst_dt, void *dst, size_t idx, float t, int round_mode) {
  switch (dst_dt) { /* jit_conv_kernel.cc:105-129 / :267-297 */
    case DFO_F32: ((float *)dst)[idx] = t; break;
    case DFO_S32: ((int32_t *)dst)[idx] = dfo_cvt_f32_s32(t, round_mode); break;
    case DFO_S8: ((int8_t *)dst)[idx] = dfo_ssat8(dfo_cvt_f32_s32(t, round_mode)); break;
    case DFO_U8: ((uint8_t *)dst)[idx] = dfo_usat8(dfo_cvt_f32_s32(t, round_mode)); break;
    default: break;
  }
}

/* conv0 output of one pixel after the fused-mode epilogue: always ReLU, round with
 * conv0_round_mode, unsigned-saturate to u8 (jit_conv_kernel.cc:256-277; defect D3 fixed
 * unless literal_f32_intermediate). */
static uint8_t conv0_requant(const dfo_conv_desc *d, int32_t acc, const void *bia0,
                             const float *scale0, int o) {
  float s = scale0[d->nscale0 > 1 ? o : 0];
  float t = dfo_epilogue_f32(acc, d->bia0_dt, bia0, o, s);
  t = dfo_relu_f32(t);
  if (d->literal_f32_intermediate && d->dst_dt == DFO_F32) {
    uint32_t bits;
    memcpy(&bits, &t, 4);
    return bits > 255u ? 255 : (uint8_t)bits;
  }
  return dfo_usat8(dfo_cvt_f32_s32(t, d->round0));
}

/* residual element `idx` (destination type) as f32: exact for u8 / s8, vcvtdq2ps for s32 */
static float load_residual_f32(int dt, const void *res, size_t idx) {
  switch (dt) {
    case DFO_F32: return ((const float *)res)[idx];
    case DFO_S32: return (float)((const int32_t *)res)[idx];
    case DFO_S8: return (float)((const int8_t *)res)[idx];
    case DFO_U8: return (float)((const uint8_t *)res)[idx];
    default: return 0.0f;
  }
}

static int conv_impl(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei,
                     const void *bia0, const float *scale0, const int8_t *wei1, const void *bia1,
                     const float *scale1, void *dst, uint8_t *mid_out, const void *res) {
  int rc = dfo_conv_check(d);
  if (rc) return rc;
  const int oh_n = dfo_conv_output_size(d->ih, d->kh, d->sh, d->ph);
  const int ow_n = dfo_conv_output_size(d->iw, d->kw, d->sw, d->pw);
  const int fused = d->oc1 > 0;
  uint8_t *mid = fused ? (uint8_t *)malloc((size_t)d->oc) : NULL;
  int8_t *w0c = (int8_t *)malloc((size_t)d->oc * d->kh * d->kw * d->ic);
  for (int o = 0; o < d->oc; ++o)
    for (int kh = 0; kh < d->kh; ++kh)
      for (int kw = 0; kw < d->kw; ++kw)
        for (int i = 0; i < d->ic; ++i)
          w0c[(((size_t)o * d->kh + kh) * d->kw + kw) * d->ic + i] =
              wei[dfo_wei_off(o, i, kh, kw, d->ic, d->kh, d->kw)];
  wei = w0c;
  /* gather w1 rows contiguously once: w1c[q*oc + o] */
  int8_t *w1c = NULL;
  if (fused && !mid_out) {
    w1c = (int8_t *)malloc((size_t)d->oc1 * d->oc);
    for (int q = 0; q < d->oc1; ++q)
      for (int o = 0; o < d->oc; ++o) w1c[(size_t)q * d->oc + o] = wei1[dfo_wei_off(q, o, 0, 0, d->oc, 1, 1)];
  }
  for (int n = 0; n < d->n; ++n)
    for (int oh = 0; oh < oh_n; ++oh)
      for (int ow = 0; ow < ow_n; ++ow) {
        size_t pix = ((size_t)n * oh_n + oh) * ow_n + ow;
        if (!fused) {
          for (int o = 0; o < d->oc; ++o) {
            int32_t acc = conv0_acc(d, src, wei, n, oh, ow, o);
            float t = dfo_epilogue_f32(acc, d->bia0_dt, bia0, o, scale0[d->nscale0 > 1 ? o : 0]);
            if (res) { /* eltwise sum: one more separately rounded add, before the ReLU */
              volatile float r = load_residual_f32(d->dst_dt, res, pix * d->oc + o), u = t;
              u = u + r;
              t = u;
            }
            if (d->relu0 || d->dst_dt == DFO_U8) t = dfo_relu_f32(t); /* :264 */
            store_dst(d->dst_dt, dst, pix * d->oc + o, t, d->round0);
          }
          continue;
        }
        for (int o = 0; o < d->oc; ++o)
          mid[o] = conv0_requant(d, conv0_acc(d, src, wei, n, oh, ow, o), bia0, scale0, o);
        if (mid_out) {
          memcpy(mid_out + pix * d->oc, mid, (size_t)d->oc);
          continue;
        }
        for (int q = 0; q < d->oc1; ++q) {
          int32_t acc1 = 0; /* C4: exact s32, partial sums over conv0 oc-chunks are exact too */
          const int8_t *wr = w1c + (size_t)q * d->oc;
          for (int o = 0; o < d->oc; ++o) acc1 += (int32_t)mid[o] * (int32_t)wr[o];
          float t = dfo_epilogue_f32(acc1, d->bia1_dt, bia1, q, scale1[d->nscale1 > 1 ? q : 0]);
          if (res) {
            volatile float r = load_residual_f32(d->dst_dt, res, pix * d->oc1 + q), u = t;
            u = u + r;
            t = u;
          }
          if (d->relu1 || d->dst_dt == DFO_U8) t = dfo_relu_f32(t); /* :102-104 */
          store_dst(d->dst_dt, dst, pix * d->oc1 + q, t, d->round1);
        }
      }
  free(mid);
  free(w1c);
  free(w0c);
  return 0;
}

int dfo_conv(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei, const void *bia0,
             const float *scale0, const int8_t *wei1, const void *bia1, const float *scale1,
             void *dst) {
  return conv_impl(d, src, wei, bia0, scale0, wei1, bia1, scale1, dst, NULL, NULL);
}

/* The operator with an eltwise sum fused into its final stage ("eltwise-sum + relu fused op", reference
 * README.md:65 -- planned there, not implemented; its yardstick is MKL-DNN's sum post-op with scale 1,
 * test/test_conv_relu_pooling.cc:118-123).  PARITY UNPINNED by the reference: these semantics are this
 * repository's definition, in the reference's own instruction style (one separately rounded vaddps between
 * vmulps and vmaxps). */
int dfo_conv_sum(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei, const void *bia0,
                 const float *scale0, const int8_t *wei1, const void *bia1, const float *scale1,
                 const void *residual, void *dst) {
  if (!residual) return -1;
  return conv_impl(d, src, wei, bia0, scale0, wei1, bia1, scale1, dst, NULL, residual);
}

int dfo_conv_intermediate(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei,
                          const void *bia0, const float *scale0, uint8_t *mid) {
  if (!d || d->oc1 <= 0) return -1;
  return conv_impl(d, src, wei, bia0, scale0, NULL, NULL, NULL, NULL, mid, NULL);
}

/* ------------------------------------------------------------------------------ pooling
 * Pooling stage of the planned "conv+relu+pooling fused op" (README.md:64).  The reference has no
 * implementation (test/test_conv_relu_pooling.cc only runs MKL-DNN: pooling_max / pooling_avg_include_padding /
 * pooling_avg_exclude_padding, zero padding, :176-225), so PARITY IS UNPINNED by the reference; this restates
 * MKL-DNN's reference pooling (simple_nhwc / ref_pooling: max over in-image elements starting from the type's
 * lowest value; avg = sum / divisor, integer types rounded from the f32 quotient). */
int dfo_pool(int dt, int kind, const void *src, void *dst, int n, int h, int w, int c, int kh, int kw, int sh,
             int sw, int ph, int pw, int oh, int ow, int round_mode) {
  if (kind < 0 || kind > 2 || n <= 0 || c <= 0 || ph >= kh || pw >= kw) return -1;
  for (int b = 0; b < n; ++b)
    for (int oy = 0; oy < oh; ++oy)
      for (int ox = 0; ox < ow; ++ox)
        for (int ch = 0; ch < c; ++ch) {
          int cnt = 0;
          long long isum = 0;
          volatile float fsum = 0.0f;
          long long imax = dt == DFO_U8 ? 0 : (dt == DFO_S8 ? -128 : INT32_MIN);
          float fmax = -FLT_MAX;
          for (int ky = 0; ky < kh; ++ky) {
            int y = oy * sh - ph + ky;
            if (y < 0 || y >= h) continue;
            for (int kx = 0; kx < kw; ++kx) {
              int x = ox * sw - pw + kx;
              if (x < 0 || x >= w) continue;
              size_t idx = (((size_t)b * h + y) * w + x) * c + ch;
              ++cnt;
              if (dt == DFO_F32) {
                float v = ((const float *)src)[idx];
                if (v > fmax) fmax = v;
                fsum = fsum + v;
              } else {
                long long v = dt == DFO_U8 ? ((const uint8_t *)src)[idx]
                                           : (dt == DFO_S8 ? ((const int8_t *)src)[idx] : ((const int32_t *)src)[idx]);
                if (v > imax) imax = v;
                isum += v;
              }
            }
          }
          size_t o = (((size_t)b * oh + oy) * ow + ox) * c + ch;
          float den = (float)(kind == 1 ? kh * kw : (cnt > 0 ? cnt : 1));
          if (dt == DFO_F32) {
            volatile float q = fsum / den;
            ((float *)dst)[o] = kind == 0 ? fmax : q;
          } else {
            long long r = imax;
            if (kind != 0) {
              volatile float q = (float)isum / den;
              r = dfo_cvt_f32_s32(q, round_mode);
            }
            if (dt == DFO_U8) ((uint8_t *)dst)[o] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            else if (dt == DFO_S8) ((int8_t *)dst)[o] = (int8_t)(r < -128 ? -128 : (r > 127 ? 127 : r));
            else ((int32_t *)dst)[o] = (int32_t)r;
          }
        }
  return 0;
}

/* ----------------------------------------------------------------------------- concat */

int dfo_concat_block(int dt, int n_inputs, const int *ic) {
  /* jit_concat_kernel.cc:157-196 */
  int ts = (dt == DFO_S8 || dt == DFO_U8) ? 1 : ((dt == DFO_F32 || dt == DFO_S32) ? 4 : 0);
  if (!ts || n_inputs <= 0) return 0;
  const int b1[] = {64, 32, 16}, b4[] = {16, 8, 4};
  const int *blocks = ts == 1 ? b1 : b4;
  int block = 0;
  for (int k = 0; k < 3; ++k) {
    block = blocks[k];
    int i = 0;
    for (; i < n_inputs; ++i)
      if (ic[i] % block) break;
    if (i == n_inputs) break;
  }
  for (int i = 0; i < n_inputs; ++i)
    if (ic[i] <= 0 || ic[i] % block) return 0;
  return block;
}

int dfo_concat(int dt, int relu, int n_inputs, const void *const *srcs, const int *ic, void *dst,
               long n_pixels) {
  if (!dfo_concat_block(dt, n_inputs, ic)) return -1;
  const int ts = (dt == DFO_S8 || dt == DFO_U8) ? 1 : 4;
  long oc = 0;
  for (int i = 0; i < n_inputs; ++i) oc += ic[i];
  for (long p = 0; p < n_pixels; ++p) { /* op_concat.cc:58-68, exactly bs*h*w pixels (D8) */
    uint8_t *out = (uint8_t *)dst + (size_t)p * oc * ts;
    for (int i = 0; i < n_inputs; ++i) {
      const uint8_t *in = (const uint8_t *)srcs[i] + (size_t)p * ic[i] * ts;
      size_t nbytes = (size_t)ic[i] * ts;
      if (!relu) {
        memcpy(out, in, nbytes);
      } else if (dt == DFO_F32) { /* vmaxps(zero, x): NaN and -0.0 pass through */
        for (int c = 0; c < ic[i]; ++c) {
          float x;
          memcpy(&x, in + 4 * c, 4);
          x = dfo_relu_f32(x);
          memcpy(out + 4 * c, &x, 4);
        }
      } else if (dt == DFO_S32) { /* vpmaxsw: signed max per 16-bit half (C6) */
        for (size_t hwd = 0; hwd < nbytes / 2; ++hwd) {
          int16_t x;
          memcpy(&x, in + 2 * hwd, 2);
          if (x < 0) x = 0;
          memcpy(out + 2 * hwd, &x, 2);
        }
      } else { /* s8 AND u8: vpmaxsb, signed max per byte (C6) */
        for (size_t b = 0; b < nbytes; ++b) out[b] = ((int8_t)in[b] < 0) ? 0 : in[b];
      }
      out += nbytes;
    }
  }
  return 0;
}
