// conv_inst_cfg4.cu -- conv_fused_kernel<GeoCfg4, dst> instantiations (compile-time geometry), see conv_kernels.cuh
#include "conv_kernels.cuh"
namespace dfconv {
KernelFn pick_static_cfg4(int dst_dt) {
  using G = GeoCfg4;
  switch (dst_dt) {
    case DF_U8: return DF_KERNEL(G, DF_U8, false, false, false);
    case DF_S8: return DF_KERNEL(G, DF_S8, false, false, false);
    case DF_S32: return DF_KERNEL(G, DF_S32, false, false, false);
    default: return DF_KERNEL(G, DF_F32, false, false, false);
  }
}
}  // namespace dfconv
