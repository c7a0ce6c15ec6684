// concat.cu -- concat(+ReLU) along channels of NHWC tensors, sm_100a.
//
// Replaces op_concat<T>::infer + jit_concat_kernel (reference src/op_concat.cc:22-72,
// src/jit_concat_kernel.cc:30-128).  The reference makes one JIT call per pixel that copies
// nb_ic[i] blocks from every input; here the whole op is one grid-stride kernel over the
// 16-byte vectors of the DESTINATION: consecutive threads write consecutive 16 B of dst (fully
// coalesced 128-bit stores) and read runs of ic[i]*sizeof(T) contiguous bytes per input that
// continue into the next pixel of the same input (sector-coalesced 128-bit loads).  HBM-bound:
// algorithmic bytes = 2 * N*H*W*sum(C)*sizeof(T).
//
// ReLU is the reference's literal one (jit_concat_kernel.cc:43-51): vpmaxsb for s8 AND u8,
// vpmaxsw for s32, vmaxps(0, x) for f32 -- see DESIGN.md (C6/D9).
#include <stdlib.h>

#include "df_common.cuh"
#include "sm100_ptx.cuh"

namespace {

constexpr int kMaxInputs = 16;  // per launch; longer lists are processed in groups
#ifndef DF_CONCAT_THREADS
#define DF_CONCAT_THREADS 256
#endif
constexpr int kThreads = DF_CONCAT_THREADS;
constexpr int kUnroll = 4;

struct ConcatParams {
  const uint4* src[kMaxInputs];
  uint32_t vec_begin[kMaxInputs + 1];  // prefix sum of 16 B vectors per pixel, per input
  uint32_t n_inputs;
  uint32_t group_vecs;     // vectors per pixel contributed by this group
  uint32_t dst_pitch;      // vectors per pixel of the whole destination
  uint32_t dst_off;        // first destination vector of this group inside a pixel
  uint32_t total;          // n_pixels * group_vecs
};

enum { kCopy = 0, kReluBytes = 1, kReluHalves = 2, kReluF32 = 3 };

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <int kMode>
__device__ __forceinline__ uint32_t relu_word(uint32_t x) {
  if (kMode == kReluBytes) return __vmaxs4(x, 0u);   // vpmaxsb
  if (kMode == kReluHalves) return __vmaxs2(x, 0u);  // vpmaxsw
  if (kMode == kReluF32) {                           // vmaxps(zero, x): NaN / -0.0 pass
    float f = __uint_as_float(x);
    return (0.0f > f) ? 0u : x;
  }
  return x;
}

// Fast path: the total thread count is a multiple of the vectors per pixel, so a thread keeps ONE
// channel position (input, offset) for the whole launch and only strides over pixels -- no division
// or input search in the loop, kUnroll independent 128-bit loads in flight.
template <int kMode>
__global__ void __launch_bounds__(kThreads) concat_kernel_strided(const ConcatParams p, uint4* __restrict__ dst,
                                                                  uint32_t n_pixels) {
  // PDL: back-to-back launches overlap this launch's scheduling and index setup with the previous
  // kernel's tail; nothing is read or written before griddep_wait()
  sm100::griddep_launch_dependents();
  const uint32_t tid = blockIdx.x * kThreads + threadIdx.x;
  const uint32_t off = tid % p.group_vecs;
  uint32_t pixel = tid / p.group_vecs;
  const uint32_t pstride = (gridDim.x * kThreads) / p.group_vecs;
  uint32_t i = 0;
  while (i + 1 < p.n_inputs && off >= p.vec_begin[i + 1]) ++i;
  const uint32_t width = p.vec_begin[i + 1] - p.vec_begin[i];
  const uint4* src = p.src[i] + (off - p.vec_begin[i]);
  uint4* out = dst + p.dst_off + off;
  sm100::griddep_wait();
  for (; pixel < n_pixels; pixel += pstride * kUnroll) {
    uint4 val[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t px = pixel + u * pstride;
      if (px < n_pixels) val[u] = ld_stream(src + (size_t)px * width);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t px = pixel + u * pstride;
      if (px < n_pixels) {
        uint4 x = val[u];
        x.x = relu_word<kMode>(x.x);
        x.y = relu_word<kMode>(x.y);
        x.z = relu_word<kMode>(x.z);
        x.w = relu_word<kMode>(x.w);
        out[(size_t)px * p.dst_pitch] = x;
      }
    }
  }
}

template <int kMode>
__global__ void __launch_bounds__(kThreads) concat_kernel(const ConcatParams p, uint4* __restrict__ dst) {
  sm100::griddep_launch_dependents();
  sm100::griddep_wait();
  const uint32_t stride = gridDim.x * kThreads;
  for (uint32_t base = blockIdx.x * kThreads + threadIdx.x; base < p.total; base += stride * kUnroll) {
    uint4 val[kUnroll];
    uint32_t dst_idx[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t v = base + u * stride;
      if (v < p.total) {
        const uint32_t pixel = v / p.group_vecs;
        const uint32_t off = v - pixel * p.group_vecs;
        uint32_t i = 0;
        while (i + 1 < p.n_inputs && off >= p.vec_begin[i + 1]) ++i;
        const uint32_t width = p.vec_begin[i + 1] - p.vec_begin[i];
        val[u] = ld_stream(p.src[i] + (size_t)pixel * width + (off - p.vec_begin[i]));
        dst_idx[u] = pixel * p.dst_pitch + p.dst_off + off;  // host guarantees < 2^32
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t v = base + u * stride;
      if (v < p.total) {
        uint4 x = val[u];
        x.x = relu_word<kMode>(x.x);
        x.y = relu_word<kMode>(x.y);
        x.z = relu_word<kMode>(x.z);
        x.w = relu_word<kMode>(x.w);
        dst[dst_idx[u]] = x;
      }
    }
  }
}

// launch with programmatic stream serialization (see the kernels)
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned blocks, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool pdl = getenv("DF_NO_PDL") == nullptr;  // read once per process, not per launch
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace

extern "C" int df_concat_check(int dtype, int n_inputs, const int* ic) {
  // jit_concat_kernel::init_conf (reference src/jit_concat_kernel.cc:130-197): 1- or 4-byte
  // dtype, and the smallest block of the candidate list must divide every input's channels.
  const int ts = df::dtype_size(dtype);
  if (!ts) return df::fail(DF_E_INVALID, "concat: unsupported dtype %d", dtype);
  if (n_inputs <= 0 || !ic) return df::fail(DF_E_INVALID, "concat: no inputs");
  const int min_block = ts == 1 ? 16 : 4;
  for (int i = 0; i < n_inputs; ++i)
    if (ic[i] <= 0 || ic[i] % min_block)
      return df::fail(DF_E_INVALID, "concat: input %d has %d channels, not a multiple of %d", i, ic[i], min_block);
  return 0;
}

extern "C" int df_concat_run(int dtype, int relu, int n_inputs, const void* const* src_dev, const int* ic,
                             void* dst_dev, long n_pixels, void* stream) {
  int rc = df_concat_check(dtype, n_inputs, ic);
  if (rc) return rc;
  if (!src_dev || !dst_dev) return df::fail(DF_E_INVALID, "concat: null device pointer");
  if (n_pixels <= 0) return 0;  // empty tensors: nothing to move
  const int ts = df::dtype_size(dtype);
  unsigned long long oc_vecs = 0;
  for (int i = 0; i < n_inputs; ++i) oc_vecs += (unsigned long long)ic[i] * ts / 16;
  if ((unsigned long long)n_pixels * oc_vecs >= (1ull << 32))
    return df::fail(DF_E_UNSUPPORTED, "concat: more than 2^32 16-byte vectors in one call");
  if ((reinterpret_cast<uintptr_t>(dst_dev) & 15))
    return df::fail(DF_E_INVALID, "concat: dst not 16-byte aligned");
  const int mode = !relu ? kCopy : (dtype == DF_F32 ? kReluF32 : (dtype == DF_S32 ? kReluHalves : kReluBytes));
  static thread_local int sms = 0;  // queried once per thread: keeps the per-launch host cost low
  if (sms <= 0 && df_device_sm_count(&sms) != 0) sms = 148;
  uint32_t done_vecs = 0;
  for (int g0 = 0; g0 < n_inputs; g0 += kMaxInputs) {
    ConcatParams p;
    p.n_inputs = (uint32_t)((n_inputs - g0) < kMaxInputs ? (n_inputs - g0) : kMaxInputs);
    p.vec_begin[0] = 0;
    for (uint32_t i = 0; i < p.n_inputs; ++i) {
      if (reinterpret_cast<uintptr_t>(src_dev[g0 + i]) & 15)
        return df::fail(DF_E_INVALID, "concat: src %d not 16-byte aligned", g0 + (int)i);
      p.src[i] = static_cast<const uint4*>(src_dev[g0 + i]);
      p.vec_begin[i + 1] = p.vec_begin[i] + (uint32_t)(ic[g0 + i] * ts / 16);
    }
    p.group_vecs = p.vec_begin[p.n_inputs];
    p.dst_pitch = (uint32_t)oc_vecs;
    p.dst_off = done_vecs;
    p.total = (uint32_t)n_pixels * p.group_vecs;
    done_vecs += p.group_vecs;
    const unsigned per_block = kThreads * kUnroll;
    unsigned blocks = (p.total + per_block - 1) / per_block;
    const unsigned cap = (unsigned)sms * (2048 / kThreads);  // a multiple of the SM count, all CTAs resident
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    // strided fast path when a block-count with (blocks * 256) % vectors-per-pixel == 0 exists nearby
    uint32_t g = p.group_vecs, a = kThreads;
    while (a) { uint32_t t = g % a; g = a; a = t; }             // g = gcd(group_vecs, 256)
    const uint32_t block_multiple = p.group_vecs / g;            // blocks must be a multiple of this
    if (block_multiple <= 64 && (uint32_t)n_pixels >= 64) {
      unsigned want = (p.total + kThreads * kUnroll - 1) / (kThreads * kUnroll);
      if (want > cap) want = cap;
      unsigned sblocks = (want + block_multiple - 1) / block_multiple * block_multiple;
      uint4* d4 = (uint4*)dst_dev;
      const uint32_t npx = (uint32_t)n_pixels;
      switch (mode) {
        case kCopy: DF_CUDA(launch_pdl(concat_kernel_strided<kCopy>, sblocks, st, p, d4, npx)); break;
        case kReluBytes: DF_CUDA(launch_pdl(concat_kernel_strided<kReluBytes>, sblocks, st, p, d4, npx)); break;
        case kReluHalves: DF_CUDA(launch_pdl(concat_kernel_strided<kReluHalves>, sblocks, st, p, d4, npx)); break;
        default: DF_CUDA(launch_pdl(concat_kernel_strided<kReluF32>, sblocks, st, p, d4, npx)); break;
      }
      continue;
    }
    uint4* d4 = (uint4*)dst_dev;
    switch (mode) {
      case kCopy: DF_CUDA(launch_pdl(concat_kernel<kCopy>, blocks, st, p, d4)); break;
      case kReluBytes: DF_CUDA(launch_pdl(concat_kernel<kReluBytes>, blocks, st, p, d4)); break;
      case kReluHalves: DF_CUDA(launch_pdl(concat_kernel<kReluHalves>, blocks, st, p, d4)); break;
      default: DF_CUDA(launch_pdl(concat_kernel<kReluF32>, blocks, st, p, d4)); break;
    }
  }
  return 0;
}
