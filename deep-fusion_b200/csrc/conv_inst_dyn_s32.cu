// conv_inst_dyn_s32.cu -- run-time-geometry conv_fused_kernel instantiations for s32 destinations
#include "conv_kernels.cuh"
namespace dfconv {
KernelFn pick_dynamic_s32(bool down0, bool down1, bool nan_safe) {
  if (nan_safe) {
    if (down0) return down1 ? DF_KERNEL(DynGeom, DF_S32, true, true, true) : DF_KERNEL(DynGeom, DF_S32, true, false, true);
    return down1 ? DF_KERNEL(DynGeom, DF_S32, false, true, true) : DF_KERNEL(DynGeom, DF_S32, false, false, true);
  }
  if (down0) return down1 ? DF_KERNEL(DynGeom, DF_S32, true, true, false) : DF_KERNEL(DynGeom, DF_S32, true, false, false);
  return down1 ? DF_KERNEL(DynGeom, DF_S32, false, true, false) : DF_KERNEL(DynGeom, DF_S32, false, false, false);
}
}  // namespace dfconv
