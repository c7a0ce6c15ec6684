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
// Destroys an op together with the device resources it owns.  Plain destruction through std::unique_ptr<op>
// cannot do that: the reference's `op` (include/deepfusion.h:105-114, kept verbatim) has no virtual destructor,
// so the resources of an op that is simply dropped are reclaimed only when its address is reused or at exit.
void release(std::unique_ptr<op> &o);

}  // namespace ext
}  // namespace deepfusion
