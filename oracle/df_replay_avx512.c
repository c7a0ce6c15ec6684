/* df_replay_avx512.c -- AVX-512-VNNI + OpenMP replay of the reference's CPU path.
 * TEST / BASELINE INFRASTRUCTURE ONLY (see df_oracle.h); never linked into the product.
 *
 * The reference generates its kernels at run time with Xbyak, which is not available here, so
 * this file issues the same instruction sequence through intrinsics, with the same blocking and
 * the same loop nest:
 *   driver  : op_conv<T>::infer_conv0conv1 (src/op_conv.cc:140-260) -- balance211 over bs*oh rows,
 *             occ -> icc -> row loops, per-thread s32 workspaces -- with the addressing defects
 *             D2 corrected;
 *   kernel  : jit_conv_kernel::generate / compute_loop / store_output / compute1x1_loop /
 *             store_1x1output (src/jit_conv_kernel.cc:27-510): ur_w x nb_oc_blocking zmm
 *             accumulators, vpbroadcastd + vpdpbusd MACs on OIhw4i16o4i weights, epilogue
 *             vcvtdq2ps / vaddps / vmulps / vmaxps / vcvtps2dq{rn,rd} / vpmovusdb, u8 tile kept
 *             in xmm registers and consumed by the 1x1 loop via vpextrd + vpbroadcastd;
 *   concat  : op_concat<T>::infer + jit_concat_kernel (src/op_concat.cc:22-72,
 *             src/jit_concat_kernel.cc:30-128): per pixel, per input, nb_ic blocks of
 *             vmovups / vpmaxs{b,w} | vmaxps / vmovups.
 * It serves two purposes: (1) a second, independent statement of the arithmetic that must agree
 * bit-for-bit with the scalar oracle, using the real x86 instructions; (2) the timed CPU
 * baseline ("kind": "port") beside the GPU numbers.
 */
#include <immintrin.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#include "df_oracle.h"

#define TGT __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx512vnni")))
#define INL static inline __attribute__((always_inline))

int dfr_supported(void) {
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
         __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vnni");
}
int dfr_num_threads(void) { return omp_get_max_threads(); }

typedef struct {
  dfo_conv_desc d;
  int oh, ow;
  int nb_ic, nb_oc, nb_oc1;
  int nb_ic_blocking, nb_oc_blocking, ur_w, ur_w_tail;
  const uint8_t *src;
  const int8_t *wei, *wei1;
  const void *bia0, *bia1;
  const float *scale0, *scale1;
  void *dst;
} ctx_t;

/* per-call arguments = jit_conv_call_t (src/jit_call_conf.h:48-65) */
typedef struct {
  const uint8_t *src; /* row ih_first of the window, channel chunk applied          */
  const int8_t *wei;  /* (occ, icc, kh_first) applied                                */
  int32_t *acc_s32;   /* conv0 partial sums for this row                             */
  int32_t *acc1x1;    /* conv1 partial sums for this row (oc1/16, ow, 16)            */
  uint8_t *dst;       /* output row                                                  */
  int kh_padding;     /* number of valid kernel rows                                 */
  int icb, ocb;       /* first 16-block of this ic / oc chunk                        */
} call_t;

TGT INL __m512 load_bias16(int dt, const void *bia, int ch0) {
  switch (dt) { /* jit_conv_kernel.cc:238-254 */
    case DFO_F32: return _mm512_loadu_ps((const float *)bia + ch0);
    case DFO_S32: return _mm512_cvtepi32_ps(_mm512_loadu_si512((const int32_t *)bia + ch0));
    case DFO_S8:
      return _mm512_cvtepi32_ps(
          _mm512_cvtepi8_epi32(_mm_loadu_si128((const __m128i *)((const int8_t *)bia + ch0))));
    case DFO_U8:
      return _mm512_cvtepi32_ps(
          _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)((const uint8_t *)bia + ch0))));
    default: return _mm512_setzero_ps();
  }
}

TGT INL __m512i cvt_round(__m512 v, int mode) {
  return mode == DFO_DOWN ? _mm512_cvt_roundps_epi32(v, _MM_FROUND_TO_NEG_INF | _MM_FROUND_NO_EXC)
                          : _mm512_cvt_roundps_epi32(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
}

/* One ur_w block of one output row: compute_loop + store_output (+ compute1x1_loop). */
TGT INL void row_block(const ctx_t *c, const call_t *p, const int UR, const int NB, int ow0) {
  const dfo_conv_desc *d = &c->d;
  const int ic = d->ic, kw_n = d->kw, sw = d->sw;
  const int first_ic = p->icb == 0;
  const int last_ic = p->icb + c->nb_ic_blocking >= c->nb_ic;
  __m512i acc[4][14];

  /* prepare_output (:193-216) */
#pragma GCC unroll 4
  for (int k = 0; k < NB; ++k)
#pragma GCC unroll 14
    for (int j = 0; j < UR; ++j)
      acc[k][j] = first_ic ? _mm512_setzero_si512()
                           : _mm512_loadu_si512(p->acc_s32 + ((size_t)ow0 * NB + k * UR + j) * 16);

  /* compute_loop (:358-389); taps left/right of the image are not executed */
  for (int kj = 0; kj < p->kh_padding; ++kj) {
    const uint8_t *in_row = p->src + (size_t)kj * d->iw * ic;
    const int8_t *w_row = p->wei + (size_t)kj * kw_n * 256;
    for (int ki = 0; ki < kw_n; ++ki) {
      int jj_start = 0, jj_end = UR;
      while (jj_start < UR && (ow0 + jj_start) * sw - d->pw + ki < 0) ++jj_start;
      while (jj_end > jj_start && (ow0 + jj_end - 1) * sw - d->pw + ki >= d->iw) --jj_end;
      if (jj_end <= jj_start) continue;
      for (int cc = 0; cc < c->nb_ic_blocking; ++cc)
        for (int i4 = 0; i4 < 4; ++i4) {
          __m512i inp[14];
#pragma GCC unroll 14
          for (int jj = 0; jj < UR; ++jj)
            if (jj >= jj_start && jj < jj_end) {
              int iw = (ow0 + jj) * sw - d->pw + ki;
              inp[jj] = _mm512_set1_epi32(*(const int32_t *)(in_row + (size_t)iw * ic + cc * 16 + i4 * 4));
            }
#pragma GCC unroll 4
          for (int ii = 0; ii < NB; ++ii) {
            /* kernel_offset (:333-338) */
            const int8_t *wp = w_row + (size_t)ii * c->nb_ic * d->kh * kw_n * 256 +
                               (size_t)cc * d->kh * kw_n * 256 + ki * 256 + i4 * 64;
            __m512i w = _mm512_loadu_si512(wp);
#pragma GCC unroll 14
            for (int jj = 0; jj < UR; ++jj)
              if (jj >= jj_start && jj < jj_end) acc[ii][jj] = _mm512_dpbusd_epi32(acc[ii][jj], inp[jj], w);
          }
        }
    }
  }

  if (!last_ic) { /* l_update_acc (:307-313) */
#pragma GCC unroll 4
    for (int k = 0; k < NB; ++k)
#pragma GCC unroll 14
      for (int j = 0; j < UR; ++j)
        _mm512_storeu_si512(p->acc_s32 + ((size_t)ow0 * NB + k * UR + j) * 16, acc[k][j]);
    return;
  }

  /* store_output (:228-300) */
  const int fused = d->oc1 > 0;
  const __m512 zero = _mm512_setzero_ps();
  __m128i mid[4][14];
#pragma GCC unroll 4
  for (int k = 0; k < NB; ++k) {
    const int ch0 = (p->ocb + k) * 16;
    const __m512 scale =
        d->nscale0 > 1 ? _mm512_loadu_ps(c->scale0 + ch0) : _mm512_set1_ps(c->scale0[0]); /* D4 */
    const __m512 bias = d->bia0_dt != DFO_UNDEF ? load_bias16(d->bia0_dt, c->bia0, ch0) : zero;
#pragma GCC unroll 14
    for (int j = 0; j < UR; ++j) {
      __m512 t = _mm512_cvtepi32_ps(acc[k][j]);
      if (d->bia0_dt != DFO_UNDEF) t = _mm512_add_ps(t, bias);
      t = _mm512_mul_ps(t, scale);
      if (d->relu0 || d->dst_dt == DFO_U8 || fused) t = _mm512_max_ps(zero, t);
      if (fused) {
        if (d->literal_f32_intermediate && d->dst_dt == DFO_F32)
          mid[k][j] = _mm512_cvtusepi32_epi8(_mm512_castps_si512(t)); /* defect D3, literal */
        else
          mid[k][j] = _mm512_cvtusepi32_epi8(cvt_round(t, d->round0));
      } else {
        uint8_t *o = p->dst + ((size_t)(ow0 + j) * d->oc + ch0) * (d->dst_dt == DFO_F32 || d->dst_dt == DFO_S32 ? 4 : 1);
        switch (d->dst_dt) {
          case DFO_F32: _mm512_storeu_ps((float *)o, t); break;
          case DFO_S32: _mm512_storeu_si512(o, cvt_round(t, d->round0)); break;
          case DFO_S8: _mm_storeu_si128((__m128i *)o, _mm512_cvtsepi32_epi8(cvt_round(t, d->round0))); break;
          default: _mm_storeu_si128((__m128i *)o, _mm512_cvtusepi32_epi8(cvt_round(t, d->round0))); break;
        }
      }
    }
  }
  if (!fused) return;

  /* compute1x1_loop (:143-191) */
  const int first_oc = p->ocb == 0;
  const int last_oc = p->ocb + c->nb_oc_blocking >= c->nb_oc;
  const size_t ts_out = (d->dst_dt == DFO_F32 || d->dst_dt == DFO_S32) ? 4 : 1;
  for (int ob1 = 0; ob1 < c->nb_oc1; ++ob1) {
    __m512i a1[14];
    int32_t *ws = p->acc1x1 + ((size_t)ob1 * c->ow + ow0) * 16; /* (oc1/16, ow, 16o) */
#pragma GCC unroll 14
    for (int j = 0; j < UR; ++j)
      a1[j] = first_oc ? _mm512_setzero_si512() : _mm512_loadu_si512(ws + (size_t)j * 16);
    const int8_t *w1 = c->wei1 + (size_t)ob1 * d->oc * 16 + (size_t)p->ocb * 256;
#pragma GCC unroll 4
    for (int k = 0; k < NB; ++k)
#pragma GCC unroll 4
      for (int i4 = 0; i4 < 4; ++i4) {
        __m512i w = _mm512_loadu_si512(w1 + (size_t)(k * 4 + i4) * 64);
#pragma GCC unroll 14
        for (int j = 0; j < UR; ++j) {
          int32_t four;
          switch (i4) { /* vmovd / vpextrd (:177-182) */
            case 0: four = _mm_cvtsi128_si32(mid[k][j]); break;
            case 1: four = _mm_extract_epi32(mid[k][j], 1); break;
            case 2: four = _mm_extract_epi32(mid[k][j], 2); break;
            default: four = _mm_extract_epi32(mid[k][j], 3); break;
          }
          a1[j] = _mm512_dpbusd_epi32(a1[j], _mm512_set1_epi32(four), w);
        }
      }
    if (!last_oc) { /* l_update_acc (:133-139) */
#pragma GCC unroll 14
      for (int j = 0; j < UR; ++j) _mm512_storeu_si512(ws + (size_t)j * 16, a1[j]);
      continue;
    }
    /* store_1x1output (:59-130) */
    const int ch0 = ob1 * 16;
    const __m512 scale =
        d->nscale1 > 1 ? _mm512_loadu_ps(c->scale1 + ch0) : _mm512_set1_ps(c->scale1[0]);
    const __m512 bias = d->bia1_dt != DFO_UNDEF ? load_bias16(d->bia1_dt, c->bia1, ch0) : zero;
#pragma GCC unroll 14
    for (int j = 0; j < UR; ++j) {
      __m512 t = _mm512_cvtepi32_ps(a1[j]);
      if (d->bia1_dt != DFO_UNDEF) t = _mm512_add_ps(t, bias);
      t = _mm512_mul_ps(t, scale);
      if (d->relu1 || d->dst_dt == DFO_U8) t = _mm512_max_ps(zero, t);
      uint8_t *o = p->dst + ((size_t)(ow0 + j) * d->oc1 + ch0) * ts_out;
      switch (d->dst_dt) {
        case DFO_F32: _mm512_storeu_ps((float *)o, t); break;
        case DFO_S32: _mm512_storeu_si512(o, cvt_round(t, d->round1)); break;
        case DFO_S8: _mm_storeu_si128((__m128i *)o, _mm512_cvtsepi32_epi8(cvt_round(t, d->round1))); break;
        default: _mm_storeu_si128((__m128i *)o, _mm512_cvtusepi32_epi8(cvt_round(t, d->round1))); break;
      }
    }
  }
}

/* compile-time (UR, NB) instances, the analogue of one JIT-generated kernel body */
#define INST(UR, NB) \
  TGT static void rb_##UR##_##NB(const ctx_t *c, const call_t *p, int ow0) { row_block(c, p, UR, NB, ow0); }
INST(1, 4) INST(2, 4) INST(3, 4) INST(4, 4) INST(5, 4)
INST(1, 3) INST(2, 3) INST(3, 3) INST(4, 3) INST(5, 3) INST(6, 3) INST(7, 3)
INST(1, 2) INST(2, 2) INST(3, 2) INST(4, 2) INST(5, 2) INST(6, 2) INST(7, 2) INST(8, 2) INST(9, 2)
INST(1, 1) INST(2, 1) INST(3, 1) INST(4, 1) INST(5, 1) INST(6, 1) INST(7, 1) INST(8, 1) INST(9, 1)
INST(10, 1) INST(11, 1) INST(12, 1) INST(13, 1) INST(14, 1)

typedef void (*rb_fn)(const ctx_t *, const call_t *, int);
static rb_fn pick(int ur, int nb) {
#define P(UR, NB) if (ur == UR && nb == NB) return rb_##UR##_##NB;
  P(1, 4) P(2, 4) P(3, 4) P(4, 4) P(5, 4)
  P(1, 3) P(2, 3) P(3, 3) P(4, 3) P(5, 3) P(6, 3) P(7, 3)
  P(1, 2) P(2, 2) P(3, 2) P(4, 2) P(5, 2) P(6, 2) P(7, 2) P(8, 2) P(9, 2)
  P(1, 1) P(2, 1) P(3, 1) P(4, 1) P(5, 1) P(6, 1) P(7, 1) P(8, 1) P(9, 1)
  P(10, 1) P(11, 1) P(12, 1) P(13, 1) P(14, 1)
#undef P
  return NULL;
}

/* generate() (:395-510): one output row as ur_w blocks plus a tail block */
static void jit_ker(const ctx_t *c, const call_t *p, rb_fn full, rb_fn tail) {
  int ow0 = 0;
  for (; ow0 + c->ur_w <= c->ow; ow0 += c->ur_w) full(c, p, ow0);
  if (c->ur_w_tail) tail(c, p, ow0);
}

int dfr_conv(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei, const void *bia0,
             const float *scale0, const int8_t *wei1, const void *bia1, const float *scale1,
             void *dst) {
  if (!dfr_supported()) return -100;
  int rc = dfo_conv_check(d);
  if (rc) return rc;
  ctx_t c;
  memset(&c, 0, sizeof c);
  c.d = *d;
  c.oh = dfo_conv_output_size(d->ih, d->kh, d->sh, d->ph);
  c.ow = dfo_conv_output_size(d->iw, d->kw, d->sw, d->pw);
  c.nb_ic = d->ic / 16;
  c.nb_oc = d->oc / 16;
  c.nb_oc1 = d->oc1 / 16;
  int blk[4];
  dfo_conv_blocking(d->ic, d->oc, c.ow, d->kh, d->kw, blk);
  c.nb_ic_blocking = blk[0];
  c.nb_oc_blocking = blk[1];
  c.ur_w = blk[2];
  c.ur_w_tail = blk[3];
  c.src = src; c.wei = wei; c.wei1 = wei1; c.bia0 = bia0; c.bia1 = bia1;
  c.scale0 = scale0; c.scale1 = scale1; c.dst = dst;
  rb_fn full = pick(c.ur_w, c.nb_oc_blocking);
  rb_fn tail = c.ur_w_tail ? pick(c.ur_w_tail, c.nb_oc_blocking) : NULL;
  if (!full || (c.ur_w_tail && !tail)) return -101;

  const int oc_chunks = c.nb_oc / c.nb_oc_blocking, ic_chunks = c.nb_ic / c.nb_ic_blocking;
  const int fused = d->oc1 > 0;
  const size_t ts_out = (d->dst_dt == DFO_F32 || d->dst_dt == DFO_S32) ? 4 : 1;
  const size_t out_c = fused ? d->oc1 : d->oc;
  const size_t ws_per_thread = (size_t)c.oh * c.ow * 16 * c.nb_oc_blocking;       /* op_conv.h:72 */
  const size_t ws1_per_thread = fused ? (size_t)c.oh * c.ow * d->oc1 : 16;        /* op_conv.h:76 */
  const int nthr_max = omp_get_max_threads();
  int32_t *ws = (int32_t *)aligned_alloc(4096, ((nthr_max * ws_per_thread * 4 + 4095) / 4096) * 4096);
  int32_t *ws1 = (int32_t *)aligned_alloc(4096, ((nthr_max * ws1_per_thread * 4 + 4095) / 4096) * 4096);
  if (!ws || !ws1) { free(ws); free(ws1); return -102; }

#pragma omp parallel
  {
    const int ithr = omp_get_thread_num(), nthr = omp_get_num_threads();
    long start, end;
    dfo_balance211((long)d->n * c.oh, nthr, ithr, &start, &end);
    int32_t *ws_l = ws + (size_t)ithr * ws_per_thread;
    int32_t *ws1_l = ws1 + (size_t)ithr * ws1_per_thread;
    while (start < end) {
      const int n = (int)(start / c.oh), oh_s = (int)(start % c.oh);
      const long work_rem = end - start;
      const int oh_e = oh_s + work_rem > c.oh ? c.oh : (int)(oh_s + work_rem);
      for (int occ = 0; occ < (fused ? oc_chunks : oc_chunks); ++occ) {
        const int ocb = occ * c.nb_oc_blocking;
        for (int icc = 0; icc < ic_chunks; ++icc) {
          const int icb = icc * c.nb_ic_blocking;
          for (int oj = oh_s; oj < oh_e; ++oj) {
            const int ij = oj * d->sh - d->ph;
            const int t_over = ij < 0 ? -ij : 0;
            const int b_over = (ij + d->kh > d->ih ? ij + d->kh : d->ih) - d->ih;
            int kh_padding = d->kh - t_over - b_over;
            if (kh_padding < 0) kh_padding = 0;
            call_t p;
            /* addressing per the blocked layouts (D2 corrected) */
            p.src = src + (((size_t)n * d->ih + (ij + t_over)) * d->iw) * d->ic + (size_t)icb * 16;
            p.wei = wei + ((size_t)ocb * c.nb_ic + icb) * d->kh * d->kw * 256 + (size_t)t_over * d->kw * 256;
            p.acc_s32 = ws_l + (size_t)(oj - oh_s) * c.ow * 16 * c.nb_oc_blocking;
            p.acc1x1 = ws1_l + (size_t)(oj - oh_s) * c.ow * (fused ? d->oc1 : 0);
            p.dst = (uint8_t *)dst + (((size_t)n * c.oh + oj) * c.ow) * out_c * ts_out;
            p.kh_padding = kh_padding;
            p.icb = icb;
            p.ocb = ocb;
            jit_ker(&c, &p, full, tail);
          }
        }
      }
      start += oh_e - oh_s;
    }
  }
  free(ws);
  free(ws1);
  return 0;
}

/* ----------------------------------------------------------------------------- concat */
TGT static void concat_pixel(int dt, int relu, int bits, int n_inputs, const uint8_t *const *src,
                             const int *nb, uint8_t *dst) {
  const int step = bits / 8;
  for (int i = 0; i < n_inputs; ++i) {
    const uint8_t *s = src[i];
    for (int b = 0; b < nb[i]; ++b, s += step, dst += step) {
      if (bits == 512) {
        __m512i v = _mm512_loadu_si512(s);
        if (relu) {
          if (dt == DFO_S32) v = _mm512_max_epi16(v, _mm512_setzero_si512());
          else if (dt == DFO_F32) v = _mm512_castps_si512(_mm512_max_ps(_mm512_setzero_ps(), _mm512_castsi512_ps(v)));
          else v = _mm512_max_epi8(v, _mm512_setzero_si512());
        }
        _mm512_storeu_si512(dst, v);
      } else if (bits == 256) {
        __m256i v = _mm256_loadu_si256((const __m256i *)s);
        if (relu) {
          if (dt == DFO_S32) v = _mm256_max_epi16(v, _mm256_setzero_si256());
          else if (dt == DFO_F32) v = _mm256_castps_si256(_mm256_max_ps(_mm256_setzero_ps(), _mm256_castsi256_ps(v)));
          else v = _mm256_max_epi8(v, _mm256_setzero_si256());
        }
        _mm256_storeu_si256((__m256i *)dst, v);
      } else {
        __m128i v = _mm_loadu_si128((const __m128i *)s);
        if (relu) {
          if (dt == DFO_S32) v = _mm_max_epi16(v, _mm_setzero_si128());
          else if (dt == DFO_F32) v = _mm_castps_si128(_mm_max_ps(_mm_setzero_ps(), _mm_castsi128_ps(v)));
          else v = _mm_max_epi8(v, _mm_setzero_si128());
        }
        _mm_storeu_si128((__m128i *)dst, v);
      }
    }
  }
}

int dfr_concat(int dt, int relu, int n_inputs, const void *const *srcs, const int *ic, void *dst,
               long n_pixels) {
  if (!dfr_supported()) return -100;
  if (n_inputs > 64) return -103;
  const int block = dfo_concat_block(dt, n_inputs, ic);
  if (!block) return -1;
  const int ts = (dt == DFO_S8 || dt == DFO_U8) ? 1 : 4;
  const int bits = 8 * ts * block;
  int nb[64];
  long oc = 0;
  for (int i = 0; i < n_inputs; ++i) {
    nb[i] = ic[i] / block;
    oc += ic[i];
  }
#pragma omp parallel
  {
    const int ithr = omp_get_thread_num(), nthr = omp_get_num_threads();
    long start, end;
    dfo_balance211(n_pixels, nthr, ithr, &start, &end);
    const uint8_t *sp[64];
    for (long p = start; p < end; ++p) {
      for (int i = 0; i < n_inputs; ++i) sp[i] = (const uint8_t *)srcs[i] + (size_t)p * ic[i] * ts;
      concat_pixel(dt, relu, bits, n_inputs, sp, nb, (uint8_t *)dst + (size_t)p * oc * ts);
    }
  }
  return 0;
}
