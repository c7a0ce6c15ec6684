/* deepfusion_c.h -- flat C view of the C++ API in deepfusion.h (+ deepfusion_ext.h).
 *
 * Exists so that non-C++ hosts (bench.py and the parity tests through ctypes; a cgo / JNI / N-API
 * binding in general) can drive exactly the call sequence a C++ user of the reference makes:
 * memory(...), concat(...)/conv(...), op->submit().  Semantics, including "creation failure prints
 * [ERROR ...] and exits the process" (reference util/log.h:38-42), are those of the C++ API.
 */
#ifndef DEEPFUSION_C_H_
#define DEEPFUSION_C_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfh_memory dfh_memory;
typedef struct dfh_op dfh_op;

/* format / dtype use the numeric values of deepfusion::memory::format / ::dtype */
dfh_memory *dfh_memory_create_nchw(const int dims_nchw[4], int format, int dtype, int alignment);
dfh_memory *dfh_memory_create(const int *dims, int ndims, int format, int dtype, int alignment);
void *dfh_memory_data(dfh_memory *m);   /* host pointer */
size_t dfh_memory_bytes(dfh_memory *m);
void *dfh_memory_device(dfh_memory *m); /* ext::device_data */
void dfh_memory_pin(dfh_memory *m);     /* ext::pin */
void dfh_memory_to_device(dfh_memory *m);
void dfh_memory_to_host(dfh_memory *m);
void dfh_memory_destroy(dfh_memory *m);

dfh_op *dfh_concat_create(dfh_memory *const *srcs, int n_srcs, dfh_memory *dst, int post_relu);
/* bia / wei1x1 / bia1x1 may be NULL; wei1x1 == NULL selects the conv-only overload */
dfh_op *dfh_conv_create(dfh_memory *src, dfh_memory *wei, dfh_memory *bia, const int stride[2], const int padding[2],
                        dfh_memory *wei1x1, dfh_memory *bia1x1, dfh_memory *dst, int conv0_relu,
                        const float *conv0_scales, int n_conv0_scales, int conv0_round_mode, int conv1_relu,
                        const float *conv1_scales, int n_conv1_scales, int conv1_round_mode);
/* ext::conv_sharded(): the same operator split over `devices` by contiguous batch slabs (no collective) */
dfh_op *dfh_conv_sharded_create(const int *devices, int n_devices, dfh_memory *src, dfh_memory *wei, dfh_memory *bia,
                                const int stride[2], const int padding[2], dfh_memory *wei1x1, dfh_memory *bia1x1,
                                dfh_memory *dst, int conv0_relu, const float *conv0_scales, int n_conv0_scales,
                                int conv0_round_mode, int conv1_relu, const float *conv1_scales, int n_conv1_scales,
                                int conv1_round_mode);
/* ext::concat_conv(): concat(+ReLU) of `srcs` fused into the conv's input load */
dfh_op *dfh_concat_conv_create(dfh_memory *const *srcs, int n_srcs, int concat_relu, dfh_memory *wei, dfh_memory *bia,
                               const int stride[2], const int padding[2], dfh_memory *wei1x1, dfh_memory *bia1x1,
                               dfh_memory *dst, int conv0_relu, const float *conv0_scales, int n_conv0_scales,
                               int conv0_round_mode, int conv1_relu, const float *conv1_scales, int n_conv1_scales,
                               int conv1_round_mode);
int dfh_concat_conv_is_fused(dfh_op *op);
/* ext::conv_pool(): conv (+ReLU) into conv_dst, pooled (kind 0 max / 1 avg incl. padding / 2 avg excl.) into pool_dst */
dfh_op *dfh_conv_pool_create(dfh_memory *src, dfh_memory *wei, dfh_memory *bia, const int stride[2], const int padding[2],
                             dfh_memory *conv_dst, dfh_memory *pool_dst, int kind, const int pool_kernel[2],
                             const int pool_stride[2], const int pool_padding[2], int conv_relu, const float *conv_scales,
                             int n_conv_scales, int conv_round_mode, int pool_round_mode);
/* ext::pool(): the pooling stage on its own */
dfh_op *dfh_pool_create(dfh_memory *src, dfh_memory *dst, int kind, const int pool_kernel[2], const int pool_stride[2],
                        const int pool_padding[2], int pool_round_mode);
/* ext::conv_sum(): conv / fused conv + eltwise sum of `residual` + ReLU */
dfh_op *dfh_conv_sum_create(dfh_memory *src, dfh_memory *wei, dfh_memory *bia, const int stride[2], const int padding[2],
                            dfh_memory *wei1x1, dfh_memory *bia1x1, dfh_memory *residual, dfh_memory *dst, int conv0_relu,
                            const float *conv0_scales, int n_conv0_scales, int conv0_round_mode, int conv1_relu,
                            const float *conv1_scales, int n_conv1_scales, int conv1_round_mode);
void dfh_sharded_upload(dfh_op *op);   /* slabs -> devices (device-resident timing) */
void dfh_sharded_sync(dfh_op *op);     /* wait for every device */
void dfh_sharded_download(dfh_op *op); /* devices -> host destination, synchronous */
void dfh_op_submit(dfh_op *op);                       /* op::submit(): H2D + kernel + D2H, synchronous */
void dfh_op_submit_device(dfh_op *op, void *stream);  /* ext::submit_device(): kernel only, async */
int dfh_op_launches(dfh_op *op);
void dfh_sync(void *stream);
void dfh_op_destroy(dfh_op *op);

#ifdef __cplusplus
}
#endif
#endif
