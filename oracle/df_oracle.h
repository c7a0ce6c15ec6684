/* df_oracle.h -- CPU oracle for the deep-fusion hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a scalar C restatement of the arithmetic the reference's JIT kernels emit
 * (reference: src/jit_conv_kernel.cc, src/jit_concat_kernel.cc, src/op_conv.cc,
 * src/op_concat.cc).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call it; the product path (deep-fusion_b200/) never does.
 *
 * PARITY PINNING STATUS
 *   conv, concat: PINNED TO THE REFERENCE EXECUTING.  oracle/_ref/libdfref.so is the reference's own
 *           generator sources (src/jit_conv_kernel.cc, src/jit_concat_kernel.cc, src/op_concat.cc,
 *           src/deepfusion.cc ...), compiled unmodified from /root/reference against a recording stand-in
 *           for the un-vendored Xbyak (oracle/xbyak_shim/) and executed instruction by instruction with the
 *           host's AVX-512 units (oracle/ref_driver.cc, recipe: oracle/Makefile target _ref).
 *           tests/test_ref_pin.py asserts this oracle == that library, bit for bit, on every fused case of
 *           tests/cases.py, the conv-only operator, 16 other windows / strides / paddings, 4 fused windows,
 *           and the reference's own concat test list (test/test_concat.cc:122-153) on its data range and
 *           on full-range data.  The reference holds no golden vectors of its own (test/test_conv.cc:63-82
 *           and benchmark/bench_conv.cc:41-44 are stubs).  Further cross-checks: an independent numpy int64
 *           model (tests/np_model.py) and oracle/df_replay_avx512.c (intrinsics replay, also the timed CPU arm).
 *   helpers: dividable_of / find_dividable pinned by test/test_misc.cc:25-36; the blocking picked by
 *           jit_conv_kernel::init_conf is compared with dfo_conv_blocking in tests/test_ref_pin.py.
 *   dfo_conv_sum, dfo_pool: PARITY UNPINNED -- the reference lists these operators as planned (README.md:64-65)
 *           and has no implementation; see the comments at their definitions.
 */
#ifndef DF_ORACLE_H_
#define DF_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numbering = deepfusion::memory::dtype (reference include/deepfusion.h:66-72) */
enum { DFO_UNDEF = 0, DFO_F32 = 1, DFO_S32 = 2, DFO_S8 = 3, DFO_U8 = 4 };
/* numbering = deepfusion::round_mode (include/deepfusion.h:46-49) */
enum { DFO_NEAREST = 0, DFO_DOWN = 1 };

typedef struct {
  int n, ih, iw;          /* batch, input height/width                                  */
  int ic, oc, oc1;        /* conv0 in/out channels; oc1 = conv1x1 out channels, 0 = none */
  int kh, kw, sh, sw, ph, pw;
  int dst_dt;             /* DFO_*                                                      */
  int bia0_dt, bia1_dt;   /* DFO_UNDEF = no bias                                        */
  int relu0, relu1;
  int round0, round1;     /* DFO_NEAREST / DFO_DOWN                                     */
  int nscale0, nscale1;   /* 1 (broadcast) or oc / oc1                                  */
  int literal_f32_intermediate; /* 1 = reproduce reference defect D3       */
} dfo_conv_desc;

/* byte offset of weight (o, i, h, w) in OIhw4i16o4i (jit_conv_kernel.cc:333-338, :384) */
size_t dfo_wei_off(int o, int i, int h, int w, int ic, int kh, int kw);
/* plain oihw -> OIhw4i16o4i */
void dfo_repack_oihw(const int8_t *oihw, int8_t *blocked, int oc, int ic, int kh, int kw);

int dfo_conv_output_size(int image, int kernel, int stride, int padding);

/* shape / dtype acceptance of op_conv::init_conf + jit_conv_kernel::init_conf with defect D1
 * fixed; returns 0 when accepted, otherwise a negative reason code. */
int dfo_conv_check(const dfo_conv_desc *d);

/* conv3x3(+ReLU)(+conv1x1(+ReLU)).  src NHWC u8; weights OIhw4i16o4i s8; dst NHWC dst_dt.
 * Returns 0 or dfo_conv_check's code. */
int dfo_conv(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei, const void *bia0,
             const float *scale0, const int8_t *wei1, const void *bia1, const float *scale1,
             void *dst);
/* the u8 intermediate (conv0 output after requantisation) for debugging kernels */
int dfo_conv_intermediate(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei,
                          const void *bia0, const float *scale0, uint8_t *mid);

/* the operator with an eltwise sum of `residual` (destination type and layout) between the scale and the ReLU:
 * the reference's planned "eltwise-sum + relu fused op" (README.md:65) -- PARITY UNPINNED, see df_oracle.c */
int dfo_conv_sum(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei, const void *bia0,
                 const float *scale0, const int8_t *wei1, const void *bia1, const float *scale1,
                 const void *residual, void *dst);
/* NHWC pooling, kind 0 = max, 1 = avg (include padding), 2 = avg (exclude padding): the pooling stage of the
 * reference's planned "conv+relu+pooling fused op" (README.md:64) -- PARITY UNPINNED, see df_oracle.c */
int dfo_pool(int dt, int kind, const void *src, void *dst, int n, int h, int w, int c, int kh, int kw, int sh,
             int sw, int ph, int pw, int oh, int ow, int round_mode);

/* block size picked by jit_concat_kernel::init_conf (:157-176), 0 if rejected */
int dfo_concat_block(int dt, int n_inputs, const int *ic);
/* concat along channels of NHWC inputs with the literal ReLU of jit_concat_kernel.cc:43-51 */
int dfo_concat(int dt, int relu, int n_inputs, const void *const *srcs, const int *ic, void *dst,
               long n_pixels);

/* blocking helpers (util/deepfusion_utils.h:116-148, :190-209) */
int dfo_dividable_of(int val, const int *divisors, int n);
int dfo_find_dividable(int val, int divisor);
void dfo_balance211(long n, int team, int tid, long *start, long *end);
/* nb_ic_blocking, nb_oc_blocking, ur_w, ur_w_tail of jit_conv_kernel::init_conf :643-655 */
void dfo_conv_blocking(int ic, int oc, int ow, int kh, int kw, int out[4]);

/* element-level pieces, exported so tests can probe the rounding rules directly */
int32_t dfo_cvt_f32_s32(float t, int round_mode);   /* vcvtps2dq {rn,rd}-sae          */
float dfo_relu_f32(float t);                         /* vmaxps(zero, t)                */
uint8_t dfo_usat8(int32_t v);                        /* vpmovusdb                      */
int8_t dfo_ssat8(int32_t v);                         /* vpmovsdb                       */
float dfo_epilogue_f32(int32_t acc, int bia_dt, const void *bia, int idx, float scale);

/* ---- oracle/df_replay_avx512.c: intrinsics replay of the emitted x86 code (+OpenMP) ---- */
int dfr_supported(void);    /* 1 when the host CPU has AVX-512 F/BW/VL/VNNI */
int dfr_num_threads(void);  /* OpenMP team size the replay will use */
int dfr_conv(const dfo_conv_desc *d, const uint8_t *src, const int8_t *wei, const void *bia0,
             const float *scale0, const int8_t *wei1, const void *bia1, const float *scale1,
             void *dst);
int dfr_concat(int dt, int relu, int n_inputs, const void *const *srcs, const int *ic, void *dst,
               long n_pixels);

#ifdef __cplusplus
}
#endif
#endif
