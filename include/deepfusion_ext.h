// deepfusion_ext.h -- additive entry points that are NOT part of the reference API.
//
// The reference's memory::data() is a host pointer and submit() is synchronous.  Benchmarks and
// pipelines that keep tensors resident in HBM need to reach the device mirror and launch without
// the host <-> device copies; those hooks live here so that deepfusion.h stays the reference's API.
#pragma once
#include "deepfusion.h"

namespace deepfusion {
namespace ext {

// Device mirror of a memory (allocated on first use on the current CUDA device).
void *device_data(memory &m);
// Explicit mirror synchronisation (asynchronous on `stream`, nullptr = default stream).
void to_device(memory &m, void *stream = nullptr);
void to_host(memory &m, void *stream = nullptr);
// Launch the op on the device mirrors only: no copies, no synchronisation.
void submit_device(op &o, void *stream = nullptr);
// Wait for `stream`.
void sync(void *stream = nullptr);
// Number of kernels the op launches per submit (bench.py's gpu_launches bookkeeping).
int launches_per_submit(op &o);
// Pin the host buffer of a memory (cudaHostRegister) so submit()'s copies run at full PCIe speed.
void pin(memory &m);
// The fused conv (or, with wei1x1 == nullptr, the conv-only operator) over the GPUs `devices` of one box:
// contiguous batch slabs of ceil(N / G) images, parameters replicated, no data-path collective (the reference
// splits the same way over threads, src/op_conv.cc:155-156).  Same arguments and checks as conv(); submit()
// uploads, computes and downloads every slab and returns when the last byte has landed.  One host thread.
std::unique_ptr<op> conv_sharded(const std::vector<int> &devices, const std::unique_ptr<memory> &src,
                                 const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
                                 std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                                 const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1,
                                 std::unique_ptr<memory> &dst, bool conv0_relu = false,
                                 std::vector<float> conv0_scales = {1.f}, round_mode conv0_round_mode = round_mode::nearest,
                                 bool conv1_relu = false, std::vector<float> conv1_scales = {1.f},
                                 round_mode conv1_round_mode = round_mode::nearest);
// concat(+ReLU) fused into the conv that consumes it: the reference's concat() followed by conv() /
// fused conv(), as one operator whose kernel loads its input halo straight from `srcs` -- the concatenated tensor
// is never written.  `srcs` are nhwc u8 memories with equal N, H, W; the remaining arguments are conv()'s.  The
// result is bit-identical to concat(srcs, tmp, concat_relu) followed by conv(tmp, ...).  Inputs whose channel
// count is a multiple of 16 but not of 32 make the op run the two kernels back to back (concat_conv_is_fused()).
std::unique_ptr<op> concat_conv(const std::vector<std::unique_ptr<memory>> &srcs, bool concat_relu,
                                const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
                                std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                                const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1,
                                std::unique_ptr<memory> &dst, bool conv0_relu = false,
                                std::vector<float> conv0_scales = {1.f}, round_mode conv0_round_mode = round_mode::nearest,
                                bool conv1_relu = false, std::vector<float> conv1_scales = {1.f},
                                round_mode conv1_round_mode = round_mode::nearest);
bool concat_conv_is_fused(op &o);
// The two operators the reference lists as planned (README.md:64-65; it ships only their MKL-DNN yardstick,
// test/test_conv_relu_pooling.cc).  Semantics: include/dfcuda.h (df_pool_run, df_conv_run_sum).
enum class pool_kind { max = 0, avg_include_padding = 1, avg_exclude_padding = 2 };
// conv (+ReLU) + pooling: conv() writing `conv_dst` (kept on the device), pooled into `pool_dst` (same data type, nhwc).
std::unique_ptr<op> conv_pool(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                              const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                              std::unique_ptr<memory> &conv_dst, std::unique_ptr<memory> &pool_dst, pool_kind kind,
                              std::array<int, 2> pool_kernel, std::array<int, 2> pool_stride, std::array<int, 2> pool_padding,
                              bool conv_relu = true, std::vector<float> conv_scales = {1.f},
                              round_mode conv_round_mode = round_mode::nearest, round_mode pool_round_mode = round_mode::nearest);
// the pooling stage on its own (any of the four data types), e.g. behind a conv_sum:
std::unique_ptr<op> pool(const std::unique_ptr<memory> &src, std::unique_ptr<memory> &dst, pool_kind kind,
                         std::array<int, 2> pool_kernel, std::array<int, 2> pool_stride, std::array<int, 2> pool_padding,
                         round_mode pool_round_mode = round_mode::nearest);
// conv() / fused conv() + eltwise sum + ReLU: `residual` (dst's dims, format, data type) is added before the ReLU.
// wei1x1 == nullptr selects the conv-only operator.
std::unique_ptr<op> conv_sum(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                             const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                             const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1,
                             const std::unique_ptr<memory> &residual, std::unique_ptr<memory> &dst, bool conv0_relu = true,
                             std::vector<float> conv0_scales = {1.f}, round_mode conv0_round_mode = round_mode::nearest,
                             bool conv1_relu = true, std::vector<float> conv1_scales = {1.f},
                             round_mode conv1_round_mode = round_mode::nearest);
// device-resident use of a sharded op (timing): upload the slabs once, submit_device() launches the kernels on
// every device, sharded_sync() waits for all of them, sharded_download() brings the result back
void sharded_upload(op &o);
void sharded_sync(op &o);
void sharded_download(op &o);
// Destroys an op together with the device resources it owns.  Plain destruction through std::unique_ptr<op>
// cannot do that: the reference's `op` (include/deepfusion.h:105-114, kept verbatim) has no virtual destructor,
// so the resources of an op that is simply dropped are reclaimed only when its address is reused or at exit.
void release(std::unique_ptr<op> &o);

}  // namespace ext
}  // namespace deepfusion
