// pool.cu -- max / average pooling of NHWC tensors, sm_100a: the second half of the "conv+relu+pooling fused
// op" the reference lists as planned (README.md:64) and specifies through its MKL-DNN yardstick
// (test/test_conv_relu_pooling.cc:176-225: pooling_max / pooling_avg_include_padding /
// pooling_avg_exclude_padding over the conv's destination type, zero padding, shape list :313-391).
//
// The conv stage is the conv-only operator of conv_fused.cu; its destination stays on the device (for the
// shapes of the reference's list it is still in the 126 MB L2 when this kernel reads it), and this kernel is
// launched behind it on the same stream with programmatic dependent launch.  One thread owns 16 bytes of
// channels of one output pixel and walks the window; loads and stores are 128-bit and coalesced along the
// channel dimension.  HBM/L2-bound: algorithmic bytes = N*(H*W + OH*OW)*C*sizeof(T).
//
// Arithmetic (the oracle restates exactly this, oracle/df_oracle.c dfo_pool):
//   max : d = lowest(T); for every in-image window element s (kh outer, kw inner): if (s > d) d = s
//   avg : sum of the in-image elements (s32 for u8 / s8, s64 for s32, sequential f32 adds for f32) divided by
//         kh*kw (include padding) or by the number of in-image elements (exclude padding); integers:
//         q = float(sum) / float(count) as one f32 division, then vcvtps2dq-style rounding (nearest-even or
//         down); f32: the quotient itself.
#include <limits.h>
#include <stdlib.h>

#include "df_common.cuh"
#include "sm100_ptx.cuh"

namespace {

constexpr int kThreads = 256;

struct PoolParams {
  const uint4* src;
  uint4* dst;
  int n, h, w, oh, ow;
  int vecs;  // 16-byte vectors per pixel
  int kh, kw, sh, sw, ph, pw;
  int round_down;
  unsigned total;  // n * oh * ow * vecs
};

enum { kMax = 0, kAvgInclude = 1, kAvgExclude = 2 };

__device__ __forceinline__ int cvt_round(float t, bool down) {
  const int q = down ? __float2int_rd(t) : __float2int_rn(t);
  return (t < 2147483648.0f) ? q : (int)0x80000000;
}

template <int kDt, int kKind>
__global__ void __launch_bounds__(kThreads) pool_kernel(const PoolParams p) {
  sm100::griddep_launch_dependents();
  sm100::griddep_wait();  // the conv stage in front of us on the stream writes p.src
  constexpr int E = (kDt == DF_F32 || kDt == DF_S32) ? 4 : 16;  // elements per 16-byte vector
  for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < p.total; i += gridDim.x * kThreads) {
    const unsigned v = i % p.vecs, px = i / p.vecs;
    const int ox = px % p.ow, oy = (px / p.ow) % p.oh, n = px / (p.ow * p.oh);
    const int y0 = oy * p.sh - p.ph, x0 = ox * p.sw - p.pw;
    int cnt = 0;
    uint32_t mx[4];
    long long acc64[kDt == DF_S32 ? 4 : 1];
    int acc32[(kDt == DF_U8 || kDt == DF_S8) ? 16 : 1];
    float accf[kDt == DF_F32 ? 4 : 1];
    if constexpr (kKind == kMax) {
      const uint32_t lowest = kDt == DF_U8 ? 0u : (kDt == DF_S8 ? 0x80808080u : (kDt == DF_S32 ? 0x80000000u : 0xFF7FFFFFu));
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e] = lowest;
    } else {
#pragma unroll
      for (int e = 0; e < (kDt == DF_S32 ? 4 : 1); ++e) acc64[e] = 0;
#pragma unroll
      for (int e = 0; e < ((kDt == DF_U8 || kDt == DF_S8) ? 16 : 1); ++e) acc32[e] = 0;
#pragma unroll
      for (int e = 0; e < (kDt == DF_F32 ? 4 : 1); ++e) accf[e] = 0.f;
    }
    for (int ky = 0; ky < p.kh; ++ky) {
      const int y = y0 + ky;
      if (y < 0 || y >= p.h) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int x = x0 + kx;
        if (x < 0 || x >= p.w) continue;
        const uint4 q = p.src[((size_t)(n * p.h + y) * p.w + x) * p.vecs + v];
        const uint32_t s[4] = {q.x, q.y, q.z, q.w};
        ++cnt;
        if constexpr (kKind == kMax) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if constexpr (kDt == DF_U8) mx[e] = __vmaxu4(mx[e], s[e]);
            else if constexpr (kDt == DF_S8) mx[e] = __vmaxs4(mx[e], s[e]);
            else if constexpr (kDt == DF_S32) mx[e] = (uint32_t)max((int)mx[e], (int)s[e]);
            else mx[e] = (__uint_as_float(s[e]) > __uint_as_float(mx[e])) ? s[e] : mx[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if constexpr (kDt == DF_S32) acc64[e] += (int)s[e];
            else if constexpr (kDt == DF_F32) accf[e] = __fadd_rn(accf[e], __uint_as_float(s[e]));
            else {
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const uint32_t byte = (s[e] >> (8 * b)) & 0xffu;
                acc32[4 * e + b] += (kDt == DF_U8) ? (int)byte : (int)(int8_t)byte;
              }
            }
          }
        }
      }
    }
    uint32_t out[4];
    if constexpr (kKind == kMax) {
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = mx[e];
    } else {
      const float den = (float)(kKind == kAvgInclude ? p.kh * p.kw : (cnt > 0 ? cnt : 1));
      const bool down = p.round_down != 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if constexpr (kDt == DF_F32) out[e] = __float_as_uint(__fdiv_rn(accf[e], den));
        else if constexpr (kDt == DF_S32) out[e] = (uint32_t)cvt_round(__fdiv_rn(__ll2float_rn(acc64[e]), den), down);
        else {
          uint32_t w = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            int r = cvt_round(__fdiv_rn((float)acc32[4 * e + b], den), down);
            r = (kDt == DF_U8) ? min(max(r, 0), 255) : min(max(r, -128), 127);
            w |= ((uint32_t)r & 0xffu) << (8 * b);
          }
          out[e] = w;
        }
      }
    }
    p.dst[i] = make_uint4(out[0], out[1], out[2], out[3]);
  }
  (void)E;
}

template <class Kernel>
cudaError_t launch_pdl(Kernel kernel, unsigned blocks, cudaStream_t st, const PoolParams& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool pdl = getenv("DF_NO_PDL") == nullptr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <int kDt>
cudaError_t launch_kind(int kind, unsigned blocks, cudaStream_t st, const PoolParams& p) {
  switch (kind) {
    case kMax: return launch_pdl(pool_kernel<kDt, kMax>, blocks, st, p);
    case kAvgInclude: return launch_pdl(pool_kernel<kDt, kAvgInclude>, blocks, st, p);
    default: return launch_pdl(pool_kernel<kDt, kAvgExclude>, blocks, st, p);
  }
}

}  // namespace

extern "C" int df_pool_check(const df_pool_desc* d) {
  if (!d) return df::fail(DF_E_INVALID, "pool: null descriptor");
  const int ts = df::dtype_size(d->dtype);
  if (!ts) return df::fail(DF_E_INVALID, "pool: unsupported dtype %d", d->dtype);
  if (d->kind < 0 || d->kind > 2) return df::fail(DF_E_INVALID, "pool: kind must be DF_POOL_MAX / _AVG_INCLUDE / _AVG_EXCLUDE");
  if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c <= 0 || d->kh <= 0 || d->kw <= 0 || d->sh <= 0 || d->sw <= 0 ||
      d->ph < 0 || d->pw < 0 || d->oh <= 0 || d->ow <= 0)
    return df::fail(DF_E_INVALID, "pool: non-positive geometry");
  if ((d->c * ts) % 16) return df::fail(DF_E_INVALID, "pool: channels must fill whole 16-byte vectors (got %d x %d B)", d->c, ts);
  if (d->ph >= d->kh || d->pw >= d->kw) return df::fail(DF_E_INVALID, "pool: padding must be smaller than the window");
  // every output window must start inside the padded image and the last one must reach the last input row / column
  // region the caller claims (mkldnn pooling_forward::desc: (h + ph + ph_r - kh) / sh + 1 == oh with 0 <= ph_r)
  if ((d->oh - 1) * d->sh - d->ph >= d->h || (d->ow - 1) * d->sw - d->pw >= d->w)
    return df::fail(DF_E_INVALID, "pool: output %dx%d has windows entirely outside the %dx%d input", d->oh, d->ow, d->h, d->w);
  if (d->round_mode != DF_ROUND_NEAREST && d->round_mode != DF_ROUND_DOWN) return df::fail(DF_E_INVALID, "pool: bad round mode");
  return 0;
}

extern "C" int df_pool_run(const df_pool_desc* d, const void* src_dev, void* dst_dev, int n, void* stream) {
  int rc = df_pool_check(d);
  if (rc) return rc;
  if (!src_dev || !dst_dev) return df::fail(DF_E_INVALID, "pool: null device pointer");
  if ((reinterpret_cast<uintptr_t>(src_dev) & 15) || (reinterpret_cast<uintptr_t>(dst_dev) & 15))
    return df::fail(DF_E_INVALID, "pool: src/dst must be 16-byte aligned");
  if (n < 0 || n > d->n) return df::fail(DF_E_INVALID, "pool: batch %d outside [0, %d]", n, d->n);
  if (n == 0) return 0;
  PoolParams p;
  p.src = static_cast<const uint4*>(src_dev);
  p.dst = static_cast<uint4*>(dst_dev);
  p.n = n;
  p.h = d->h;
  p.w = d->w;
  p.oh = d->oh;
  p.ow = d->ow;
  p.vecs = d->c * df::dtype_size(d->dtype) / 16;
  p.kh = d->kh;
  p.kw = d->kw;
  p.sh = d->sh;
  p.sw = d->sw;
  p.ph = d->ph;
  p.pw = d->pw;
  p.round_down = d->round_mode == DF_ROUND_DOWN;
  const unsigned long long total = (unsigned long long)n * d->oh * d->ow * p.vecs;
  if (total >= (1ull << 32)) return df::fail(DF_E_UNSUPPORTED, "pool: more than 2^32 16-byte vectors in one call");
  p.total = (unsigned)total;
  static thread_local int sms = 0;
  if (sms <= 0 && df_device_sm_count(&sms) != 0) sms = 148;
  unsigned blocks = (p.total + kThreads - 1) / kThreads;
  const unsigned cap = (unsigned)sms * (2048 / kThreads);  // a multiple of the SM count, all CTAs resident
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d->dtype) {
    case DF_U8: DF_CUDA(launch_kind<DF_U8>(d->kind, blocks, st, p)); break;
    case DF_S8: DF_CUDA(launch_kind<DF_S8>(d->kind, blocks, st, p)); break;
    case DF_S32: DF_CUDA(launch_kind<DF_S32>(d->kind, blocks, st, p)); break;
    default: DF_CUDA(launch_kind<DF_F32>(d->kind, blocks, st, p)); break;
  }
  return 0;
}
