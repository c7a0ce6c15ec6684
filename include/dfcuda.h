/* dfcuda.h -- C-ABI of the B200 (sm_100a) implementation of deep-fusion's hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, caller-owned device pointers, no torch
 * or C++ types.  The C++ host layer (deep-fusion_b200/host/deepfusion.cc, which implements the
 * reference's public API include/deepfusion.h) is its only in-tree caller; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.  Every entry point names the reference
 * interface it stands in for (paths relative to the reference repository).
 *
 * Conventions: every function returns 0 on success, a positive cudaError_t, or a negative
 * DF_E_* code; none of them ever exits the process (the reference's error_and_exit behaviour,
 * util/log.h:38-42, is reproduced by the host layer, not here).  df_last_error() describes the
 * most recent failure on the calling thread.  There is no CPU fallback: without a CUDA device
 * the compute entry points fail with the CUDA error.
 */
#ifndef DFCUDA_H_
#define DFCUDA_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* = deepfusion::memory::dtype (include/deepfusion.h:66-72) */
enum { DF_UNDEF = 0, DF_F32 = 1, DF_S32 = 2, DF_S8 = 3, DF_U8 = 4 };
/* = deepfusion::round_mode (include/deepfusion.h:46-49) */
enum { DF_ROUND_NEAREST = 0, DF_ROUND_DOWN = 1 };

enum {
  DF_E_INVALID = -1,     /* bad argument / rejected by the reference's own init_conf rules   */
  DF_E_UNSUPPORTED = -2, /* accepted by the reference but outside the B200 path */
  DF_E_NODRIVER = -3,    /* cuTensorMapEncodeTiled not obtainable                           */
  DF_E_INTERNAL = -4
};

const char *df_last_error(void);
const char *df_version(void);

/* ---- device / memory / stream plumbing (replaces utils::aligned_malloc + implicit host
 *      residency of memory::data(), util/memory.cc:21-40, src/deepfusion.cc:76-80) ---------- */
int df_device_count(int *count);
int df_set_device(int device);
int df_get_device(int *device);
int df_device_sm_count(int *sms);
int df_malloc(size_t bytes, void **dev_ptr);
int df_free(void *dev_ptr);
int df_memset(void *dev_ptr, int value, size_t bytes, void *stream);
int df_host_register(void *host_ptr, size_t bytes); /* pin an existing host buffer */
int df_host_unregister(void *host_ptr);
int df_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int df_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int df_stream_create(void **stream);
int df_stream_sync(void *stream);
int df_stream_destroy(void *stream);
int df_event_create(void **event);
int df_event_record(void *event, void *stream);
/* inside a df_graph_begin / df_graph_end capture: records the event as a node of the graph (cudaEventRecordExternal), so that
 * every replay time-stamps it and df_event_elapsed_ms works on it; outside a capture the same as df_event_record */
int df_event_record_node(void *event, void *stream);
int df_stream_wait_event(void *stream, void *event);
int df_event_elapsed_ms(void *start, void *stop, float *ms); /* synchronises on `stop` */
int df_event_destroy(void *event);
/* CUDA graphs (no reference counterpart; the reference's callers loop over submit() on the host): capture
 * the df_* calls issued on `stream` between begin and end, then replay them with one launch each time. */
int df_graph_begin(void *stream);                    /* stream from df_stream_create, not NULL */
int df_graph_end(void *stream, void **graph_exec);
int df_graph_launch(void *graph_exec, void *stream);
int df_graph_destroy(void *graph_exec);

/* ---- concat(+ReLU): replaces op_concat<T>::infer + jit_concat_kernel
 *      (src/op_concat.cc:22-72, src/jit_concat_kernel.cc:30-197) ---------------------------
 * NHWC inputs concatenated along channels; `ic[i]` channels each; n_pixels = N*H*W.
 * Acceptance = jit_concat_kernel::init_conf: every ic[i] a multiple of 16 (1-byte dtypes) or
 * 4 (4-byte dtypes).  ReLU is the reference's literal one (signed max per byte / 16-bit half,
 * see DESIGN.md C6).  Device pointers must be 16-byte aligned. */
int df_concat_check(int dtype, int n_inputs, const int *ic);
int df_concat_run(int dtype, int relu, int n_inputs, const void *const *src_dev, const int *ic,
                  void *dst_dev, long n_pixels, void *stream);

/* ---- fused conv3x3+ReLU+conv1x1+ReLU: replaces op_conv<T> (src/op_conv.h:34-96),
 *      op_conv<T>::init_conf (src/op_conv.cc:262-365), jit_conv_kernel::init_conf
 *      (src/jit_conv_kernel.cc:512-673) and infer_conv0conv1 (src/op_conv.cc:140-260) -------- */
typedef struct df_conv_desc {
  int n, ih, iw;        /* created (maximum) batch, input height / width                    */
  int ic, oc, oc1;      /* conv0 in / out channels, conv1x1 out channels (0 = conv0 only)   */
  int kh, kw, sh, sw, ph, pw;
  int dst_dt;           /* DF_F32 / DF_S32 / DF_S8 / DF_U8                                  */
  int bia0_dt, bia1_dt; /* DF_UNDEF = no bias                                               */
  int relu0, relu1;
  int round0, round1;
  int nscale0, nscale1; /* 1 or oc / oc1                                                    */
} df_conv_desc;

typedef struct df_conv df_conv; /* opaque */

typedef struct df_conv_info {
  int tiles_per_launch;   /* 128-position tiles for the created batch                      */
  int grid, block;        /* persistent CTAs, threads per CTA                              */
  int smem_bytes;         /* dynamic shared memory per CTA                                 */
  int w0_resident, w1_resident, a_stages, b_stages;
  int padded_w, padded_h; /* Wp, Hp of the linearised padded pixel space                   */
  double macs_per_image;  /* H*W*(9*IC*OC + OC*OC1)                                        */
  double mma_efficiency;  /* real pixels / computed tile rows                              */
} df_conv_info;

/* Validates like the reference (with its defect D1 fixed), re-lays the OIhw4i16o4i weights out
 * for tcgen05, expands scales, converts biases to f32 exactly as vcvtdq2ps would, uploads to the
 * current device.  All pointers are HOST pointers and may be freed after the call. */
int df_conv_create(const df_conv_desc *desc, const int8_t *wei_OIhw4i16o4i,
                   const int8_t *wei1x1_OIhw4i16o4i, const void *bia0, const void *bia1,
                   const float *scale0, const float *scale1, df_conv **out);
/* One forward pass over `n` (<= created n) images; src NHWC u8, dst NHWC dst_dt, both device
 * pointers, 16-byte aligned.  Asynchronous on `stream`.  The current device must be the one the
 * handle was created on (DF_E_INVALID otherwise).  A handle caches tensor maps per (pointer, batch)
 * and is NOT thread-safe: use one handle per host thread (the reference has the same rule for one
 * op, src/op_conv.cc:159-160). */
int df_conv_run(df_conv *op, const uint8_t *src_dev, void *dst_dev, int n, void *stream);
/* ---- concat(+ReLU) fused into the conv's A-operand load: the producer -> consumer chain
 *      op_concat<T>::infer (src/op_concat.cc:22-72) -> op_conv<T>::infer (src/op_conv.cc:140-260) as ONE
 *      kernel.  The conv's source is the channel concatenation of `n_src` NHWC u8 tensors (src_ic[i] channels
 *      each, sum == desc->ic) which is never materialised: each K-block of the halo is loaded by TMA from the
 *      input that owns those channels.  concat_relu = the reference's literal u8 ReLU (vpmaxsb, bytes >= 128
 *      become 0), applied in shared memory.  Result == df_concat_run followed by df_conv_run, bit for bit.
 *      DF_E_UNSUPPORTED when an input's channel count is not a multiple of 32 (run the two ops instead). ---- */
int df_conv_create_concat(const df_conv_desc *desc, int n_src, const int *src_ic, int concat_relu,
                          const int8_t *wei_OIhw4i16o4i, const int8_t *wei1x1_OIhw4i16o4i, const void *bia0,
                          const void *bia1, const float *scale0, const float *scale1, df_conv **out);
int df_conv_run_concat(df_conv *op, const void *const *src_dev, void *dst_dev, int n, void *stream);
/* ---- eltwise-sum (+ReLU) fused into the operator's final stage: the "eltwise-sum + relu fused op" the reference
 *      lists as planned (README.md:65; its yardstick is MKL-DNN's sum post-op, test/test_conv_relu_pooling.cc:118-123,
 *      :148-151).  `residual_dev` has the destination's type and layout; per element
 *          t = (float(acc) + bias) * scale;  t = t + float(residual);  ReLU / round / saturate as df_conv_run,
 *      every step a separately rounded f32 operation.  Works for the conv-only and the fused operator. ---- */
int df_conv_create_sum(const df_conv_desc *desc, const int8_t *wei_OIhw4i16o4i, const int8_t *wei1x1_OIhw4i16o4i,
                       const void *bia0, const void *bia1, const float *scale0, const float *scale1, df_conv **out);
int df_conv_run_sum(df_conv *op, const uint8_t *src_dev, const void *residual_dev, void *dst_dev, int n, void *stream);
int df_conv_query(const df_conv *op, df_conv_info *info);
int df_conv_destroy(df_conv *op);
/* Diagnostic only (no reference counterpart): per-role clock64 timeline of the next launches
 * into dev_buf[grid * 4 * cap] (u64 words: tag << 48 | clock); NULL switches it off. */
int df_conv_debug_trace(df_conv *op, void *dev_buf, int cap);

/* ---- pooling stage of the "conv+relu+pooling fused op" the reference lists as planned (README.md:64; specified by
 *      its MKL-DNN yardstick, test/test_conv_relu_pooling.cc:176-225, shapes :313-391): NHWC, zero padding, in the
 *      conv's destination type.  max: largest in-image element of the window.  avg: sum of the in-image elements
 *      over kh*kw (include padding) or over their count (exclude padding); integer types round the f32 quotient
 *      with `round_mode`.  Launched behind the conv on the same stream, it reads the conv's output from L2. ---- */
enum { DF_POOL_MAX = 0, DF_POOL_AVG_INCLUDE = 1, DF_POOL_AVG_EXCLUDE = 2 };
typedef struct df_pool_desc {
  int dtype;          /* DF_F32 / DF_S32 / DF_S8 / DF_U8                                   */
  int kind;           /* DF_POOL_*                                                         */
  int n, h, w, c;     /* input (maximum) batch, height, width, channels                    */
  int kh, kw, sh, sw, ph, pw;
  int oh, ow;         /* output size; windows may run past the bottom / right edge         */
  int round_mode;     /* DF_ROUND_* for integer averages                                   */
} df_pool_desc;
int df_pool_check(const df_pool_desc *desc);
int df_pool_run(const df_pool_desc *desc, const void *src_dev, void *dst_dev, int n, void *stream);

/* ---- format tooling (host memory, no device work): the reference consumes OIhw4i16o4i /
 *      gOIhw4i16o4i weights (include/deepfusion.h:53-61, layout = jit_conv_kernel.cc:333-338) and
 *      nhwc activations but ships no converter (its tests fill blocked weights with random bytes,
 *      test/test_conv_relu_pooling.cc:251-254).  oc / ic are PER GROUP and multiples of 16. -------- */
size_t df_wei_blocked_offset(int o, int i, int h, int w, int ic, int kh, int kw);
int df_repack_oihw_to_blocked(const int8_t *oihw, int8_t *blocked, int oc, int ic, int kh, int kw);
int df_repack_blocked_to_oihw(const int8_t *blocked, int8_t *oihw, int oc, int ic, int kh, int kw);
int df_repack_goihw_to_blocked(const int8_t *goihw, int8_t *blocked, int groups, int oc, int ic, int kh, int kw);
int df_repack_blocked_to_goihw(const int8_t *blocked, int8_t *goihw, int groups, int oc, int ic, int kh, int kw);
int df_nchw_to_nhwc(const void *nchw, void *nhwc, int n, int c, int h, int w, int elem_bytes);
int df_nhwc_to_nchw(const void *nhwc, void *nchw, int n, int c, int h, int w, int elem_bytes);

#ifdef __cplusplus
}
#endif
#endif
