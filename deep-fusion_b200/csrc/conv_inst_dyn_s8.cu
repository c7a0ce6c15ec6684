// conv_inst_dyn_s8.cu -- run-time-geometry conv_fused_kernel instantiations for s8 destinations
#include "conv_kernels.cuh"
namespace dfconv {
KernelFn pick_dynamic_s8(bool down0, bool down1, bool nan_safe) {
  if (nan_safe) {
    if (down0) return down1 ? DF_KERNEL(DynGeom, DF_S8, true, true, true) : DF_KERNEL(DynGeom, DF_S8, true, false, true);
    return down1 ? DF_KERNEL(DynGeom, DF_S8, false, true, true) : DF_KERNEL(DynGeom, DF_S8, false, false, true);
  }
  if (down0) return down1 ? DF_KERNEL(DynGeom, DF_S8, true, true, false) : DF_KERNEL(DynGeom, DF_S8, true, false, false);
  return down1 ? DF_KERNEL(DynGeom, DF_S8, false, true, false) : DF_KERNEL(DynGeom, DF_S8, false, false, false);
}
}  // namespace dfconv
