// conv_inst_pair_dyn.cu -- conv_pair_kernel<DynGeom, dst>: run-time geometry on CTA pairs with streamed weight halves
// (round-to-nearest, finite constants), see conv_kernels.cuh
#include "conv_kernels.cuh"
namespace dfconv {
KernelFn pick_pair_dyn(int dst_dt) {
  using G = DynGeom;
  switch (dst_dt) {
    case DF_U8: return KernelFn{launch_pair<G, DF_U8>, attr_pair<G, DF_U8>};
    case DF_S8: return KernelFn{launch_pair<G, DF_S8>, attr_pair<G, DF_S8>};
    case DF_S32: return KernelFn{launch_pair<G, DF_S32>, attr_pair<G, DF_S32>};
    default: return KernelFn{launch_pair<G, DF_F32>, attr_pair<G, DF_F32>};
  }
}
}  // namespace dfconv
