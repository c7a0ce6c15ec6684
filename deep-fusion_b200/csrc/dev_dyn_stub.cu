// dev_dyn_stub.cu -- development builds only (make DF_MINI=1): stands in for the run-time-geometry
// instantiations so that experiments on the static kernels rebuild in seconds.  Never part of lib/.
#include "conv_kernels.cuh"
namespace dfconv {
KernelFn pick_dynamic_u8(bool, bool, bool) { return KernelFn{nullptr, nullptr}; }
KernelFn pick_dynamic_s8(bool, bool, bool) { return KernelFn{nullptr, nullptr}; }
KernelFn pick_dynamic_s32(bool, bool, bool) { return KernelFn{nullptr, nullptr}; }
KernelFn pick_dynamic_f32(bool, bool, bool) { return KernelFn{nullptr, nullptr}; }
KernelFn pick_pair_dyn(int) { return KernelFn{nullptr, nullptr}; }
}  // namespace dfconv
