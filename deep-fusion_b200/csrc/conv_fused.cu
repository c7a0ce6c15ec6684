// conv_fused.cu -- fused u8 x s8 conv3x3(s1,p1)+ReLU+conv1x1+ReLU for sm_100a (B200).
//
// Replaces op_conv<T>::infer_conv0conv1 + jit_conv_kernel (reference src/op_conv.cc:140-260,
// src/jit_conv_kernel.cc:27-510).  Same arithmetic contract, different
// machine:
//
//  * implicit GEMM on tcgen05 (kind::i8, u8 x s8 -> s32 in TMEM).  The M dimension is a
//    LINEARISED PADDED pixel space: rows of Wp >= W+1 positions (the extra columns are zero
//    padding shared by neighbouring rows), Hp = H+1 rows per image (one zero row shared by
//    neighbouring images).  In that space every 3x3 tap is a constant offset, so ONE halo
//    buffer per 128-position tile serves all nine taps: the A-operand descriptor of tap
//    (kh,kw) is the same buffer with its start address advanced by (kh*Wp+kw) rows.  The
//    hardware applies the 128/64/32-byte swizzle on absolute shared-memory address bits, so a
//    start address that is not a multiple of 8 rows needs no base_offset (probe/umma_probe.cu,
//    profiles/r01_probe.log).
//  * halo rows come in by TMA (4-D NHWC tensor map, box = {K-block, Wp, 1, 1}); out-of-image
//    rows / columns / images are zero-filled by the TMA unit -- that IS the padding.
//  * conv0 accumulates in TMEM; 8 epilogue warps apply (float(acc)+bias)*scale -> ReLU -> round
//    -> u8 and write the tile straight into shared memory in the swizzled K-major layout the
//    second GEMM wants, so the intermediate never leaves the SM.
//  * conv1x1 runs as N-chunks of <=128 output channels through two TMEM accumulators, so the
//    epilogue of chunk j overlaps the MMA of chunk j+1 and the conv0 MMAs of the next tile.
//  * warp roles: w0 TMA(A) | w1 MMA issue | w2 TMA(weights) | w3 TMEM alloc | w4-11 epilogue.
//    All hand-offs are mbarriers; tcgen05.commit releases stages.
//  * weights live in shared memory for the whole (persistent) kernel when they fit; otherwise
//    they stream through a ring of stages in exactly the order the MMA thread consumes them.
#include <math.h>
#include <string.h>

#include <vector>

#include "df_common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kTileM = 128;
constexpr int kMaxAStages = 4;
constexpr int kMaxBStages = 8;
constexpr int kAcc1Col = 256;   // TMEM column of the first conv1 accumulator
constexpr int kAcc1Stride = 128;
constexpr uint32_t kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct Params {
  int N, H, W, IC, OC, OC1;
  int Hp, Wp, NR;
  int n_tiles;
  int swb, nkb, ks_last;     // conv0: K-block bytes (= swizzle span), blocks, 32 B steps in last
  int swb1, nkb1, ks1_last;  // conv1
  int nc1, n_chunks, n_acc0;
  int SA, SB, NM, w0_res, w1_res;  // halo stages, weight stages, intermediate buffers
  uint32_t off_bias0, off_scale0, off_bias1, off_scale1;
  uint32_t off_a, a_stage_bytes, a_kb_stride;
  uint32_t off_mid, mid_bytes, mid_kb_stride;
  uint32_t off_w0, w0_block_bytes, off_w1, w1_block_bytes;
  uint32_t off_b, b_stage_bytes;
  int relu1, round0, round1, nan_safe;
  const float *bias0, *scale0, *bias1, *scale1;
  void* dst;
  unsigned long long* trace;  // optional timeline buffer (df_conv_debug_trace), normally null
  int trace_cap;
};

// Diagnostic timeline: role r of CTA b appends (tag << 48 | clock) words to its own lane of the
// buffer.  One predictable branch per event when disabled.
struct Tracer {
  unsigned long long* base;
  int cap, n;
  __device__ Tracer(const Params& p, int role) : base(nullptr), cap(p.trace_cap), n(0) {
    if (p.trace) base = p.trace + ((size_t)blockIdx.x * 4 + role) * p.trace_cap;
  }
  __device__ __forceinline__ void ev(unsigned tag) {
    if (base && n < cap) base[n++] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
  }
};

struct Barriers {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  uint64_t res_full;
  uint64_t acc0_full[2], acc0_empty[2];
  uint64_t mid_full[2], mid_empty[2];
  uint64_t acc1_full[2], acc1_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t layout_of(int swb) {
  return swb == 128 ? kLayoutSW128 : (swb == 64 ? kLayoutSW64 : kLayoutSW32);
}

// ---------------------------------------------------------------------------- epilogue math
// (float(acc) + bias) * scale as three separately rounded f32 operations -- vcvtdq2ps, vaddps,
// vmulps (jit_conv_kernel.cc:96-100, :259-263).  Never an FMA.
__device__ __forceinline__ float scale_acc(uint32_t acc, float bias, float scale) {
  return __fmul_rn(__fadd_rn(__int2float_rn((int)acc), bias), scale);
}
// vmaxps(zero, t): second source when NaN or both zero
__device__ __forceinline__ float relu_x86(float t) { return (0.0f > t) ? 0.0f : t; }

// vcvtps2dq with x86 "integer indefinite" (0x80000000) on NaN / overflow
template <bool kDown>
__device__ __forceinline__ int cvt_x86(float t) {
  int q = kDown ? __float2int_rd(t) : __float2int_rn(t);
  return (t < 2147483648.0f) ? q : (int)0x80000000;  // false for NaN and t >= 2^31
}

// ReLU -> round -> vpmovusdb for four values, packed little-endian.  Saturating a signed s32 to
// [0,255] equals ReLU followed by unsigned saturation for every finite t; NaN (only reachable
// through non-finite scales / biases) is patched to 255 when kNanSafe.
template <bool kDown, bool kNanSafe>
__device__ __forceinline__ uint32_t requant_u8x4(const uint32_t* acc, const float4 b, const float4 s) {
  float t0 = scale_acc(acc[0], b.x, s.x), t1 = scale_acc(acc[1], b.y, s.y);
  float t2 = scale_acc(acc[2], b.z, s.z), t3 = scale_acc(acc[3], b.w, s.w);
  int q0 = kDown ? __float2int_rd(t0) : __float2int_rn(t0);
  int q1 = kDown ? __float2int_rd(t1) : __float2int_rn(t1);
  int q2 = kDown ? __float2int_rd(t2) : __float2int_rn(t2);
  int q3 = kDown ? __float2int_rd(t3) : __float2int_rn(t3);
  if (kNanSafe) {
    if (t0 != t0) q0 = 255;
    if (t1 != t1) q1 = 255;
    if (t2 != t2) q2 = 255;
    if (t3 != t3) q3 = 255;
  }
  uint32_t hi, lo;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q3), "r"(q2));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(q1), "r"(q0), "r"(hi));
  return lo;
}

template <bool kDown>
__device__ __forceinline__ uint32_t requant_s8x4(const uint32_t* acc, const float4 b, const float4 s, bool relu) {
  float t[4] = {scale_acc(acc[0], b.x, s.x), scale_acc(acc[1], b.y, s.y), scale_acc(acc[2], b.z, s.z),
                scale_acc(acc[3], b.w, s.w)};
  int q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (relu) t[i] = relu_x86(t[i]);
    q[i] = cvt_x86<kDown>(t[i]);
  }
  uint32_t hi, lo;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q[3]), "r"(q[2]));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(q[1]), "r"(q[0]), "r"(hi));
  return lo;
}

// 16 accumulator columns of one row -> destination (conv1 epilogue, jit_conv_kernel.cc:89-130)
template <int kDst, bool kDown, bool kNanSafe>
__device__ __forceinline__ void store16(const uint32_t* acc, const float* bias, const float* scale, bool relu,
                                        uint8_t* out) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  const float4* s4 = reinterpret_cast<const float4*>(scale);
  if (kDst == DF_U8) {
    uint4 v;
    v.x = requant_u8x4<kDown, kNanSafe>(acc + 0, b4[0], s4[0]);
    v.y = requant_u8x4<kDown, kNanSafe>(acc + 4, b4[1], s4[1]);
    v.z = requant_u8x4<kDown, kNanSafe>(acc + 8, b4[2], s4[2]);
    v.w = requant_u8x4<kDown, kNanSafe>(acc + 12, b4[3], s4[3]);
    *reinterpret_cast<uint4*>(out) = v;
  } else if (kDst == DF_S8) {
    uint4 v;
    v.x = requant_s8x4<kDown>(acc + 0, b4[0], s4[0], relu);
    v.y = requant_s8x4<kDown>(acc + 4, b4[1], s4[1], relu);
    v.z = requant_s8x4<kDown>(acc + 8, b4[2], s4[2], relu);
    v.w = requant_s8x4<kDown>(acc + 12, b4[3], s4[3], relu);
    *reinterpret_cast<uint4*>(out) = v;
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 b = b4[g], s = s4[g];
      float t0 = scale_acc(acc[4 * g + 0], b.x, s.x), t1 = scale_acc(acc[4 * g + 1], b.y, s.y);
      float t2 = scale_acc(acc[4 * g + 2], b.z, s.z), t3 = scale_acc(acc[4 * g + 3], b.w, s.w);
      if (relu) {
        t0 = relu_x86(t0);
        t1 = relu_x86(t1);
        t2 = relu_x86(t2);
        t3 = relu_x86(t3);
      }
      uint4 v;
      if (kDst == DF_F32) {
        v = make_uint4(__float_as_uint(t0), __float_as_uint(t1), __float_as_uint(t2), __float_as_uint(t3));
      } else {
        v = make_uint4((uint32_t)cvt_x86<kDown>(t0), (uint32_t)cvt_x86<kDown>(t1), (uint32_t)cvt_x86<kDown>(t2),
                       (uint32_t)cvt_x86<kDown>(t3));
      }
      reinterpret_cast<uint4*>(out)[g] = v;
    }
  }
}

// ------------------------------------------------------------------------------- the kernel
template <int kDst, bool kDown0, bool kDown1, bool kNanSafe>
__global__ void __launch_bounds__(kThreads, 1)
conv_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
                  const __grid_constant__ CUtensorMap tmW1, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align in the shared address space (keeps LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Barriers* bar = reinterpret_cast<Barriers*>(smem);
  const uint32_t sbase = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(smem_u32(&bar->a_full[i]), 1);
      mbar_init(smem_u32(&bar->a_empty[i]), 1);
    }
    for (int i = 0; i < kMaxBStages; ++i) {
      mbar_init(smem_u32(&bar->b_full[i]), 1);
      mbar_init(smem_u32(&bar->b_empty[i]), 1);
    }
    mbar_init(smem_u32(&bar->res_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar->acc0_full[i]), 1);
      mbar_init(smem_u32(&bar->acc0_empty[i]), kEpiWarps);
      mbar_init(smem_u32(&bar->mid_full[i]), kEpiWarps);
      mbar_init(smem_u32(&bar->mid_empty[i]), 1);
      mbar_init(smem_u32(&bar->acc1_full[i]), 1);
      mbar_init(smem_u32(&bar->acc1_empty[i]), kEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
  }
  if (warp == 3) tmem_alloc<512>(smem_u32(&bar->tmem_base));
  {
    // per-channel f32 bias / scale vectors -> smem (read by every epilogue thread)
    float* sb0 = reinterpret_cast<float*>(smem + p.off_bias0);
    float* ss0 = reinterpret_cast<float*>(smem + p.off_scale0);
    float* sb1 = reinterpret_cast<float*>(smem + p.off_bias1);
    float* ss1 = reinterpret_cast<float*>(smem + p.off_scale1);
    for (int i = threadIdx.x; i < p.OC; i += kThreads) {
      sb0[i] = p.bias0[i];
      ss0[i] = p.scale0[i];
    }
    const int oc1_pad = p.n_chunks * p.nc1;
    for (int i = threadIdx.x; i < oc1_pad; i += kThreads) {
      sb1[i] = i < p.OC1 ? p.bias1[i] : 0.f;
      ss1[i] = i < p.OC1 ? p.scale1[i] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bar->tmem_base;

  const int q_first = 2 * p.Wp;  // linear index of image 0, row 0, column 0

  // The three single-thread roles each run their whole loop inside ONE elect.sync region and advance
  // shared-memory descriptors by ADDITION: that keeps descriptor math in the uniform datapath.  Both
  // alternatives measured slow (probe/mma_contention.cu, profiles/r01_mma_issue_probe.log): a
  // per-tap elect/__syncwarp costs ~370 cycles per iteration, and rebuilding descriptors from
  // vector registers (R2UR) ~140 cycles per tap -- more than the 96..256 cycles of MMA work in a tap.
  if (warp == 0) {
    // =============================== TMA producer: halo rows ===============================
    if (elect_one()) {
      Tracer tr(p, 0);
      for (int it = 0; it < n_local; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int s = it % p.SA;
        mbar_wait(smem_u32(&bar->a_empty[s]), ((it / p.SA) & 1) ^ 1);
        tr.ev(1);
        const int q0 = q_first + tile * kTileM;
        const int g_lo = (q0 - p.Wp - 1) / p.Wp;
        const int g_hi = (q0 + kTileM + p.Wp) / p.Wp;
        const int nrows = g_hi - g_lo + 1;
        const uint32_t full = smem_u32(&bar->a_full[s]);
        const uint32_t stage = sbase + p.off_a + s * p.a_stage_bytes;
        mbar_expect_tx(full, (uint32_t)(nrows * p.nkb * p.Wp * p.swb));
        int n = (g_lo > 0) ? (g_lo - 1) / p.Hp : 0;
        int h = (g_lo > 0) ? (g_lo - 1) - n * p.Hp - 1 : -2;  // -2: the all-zero row above everything
        uint32_t dst = stage;
        const uint32_t row_bytes = p.Wp * p.swb;
        for (int r = 0; r < nrows; ++r, dst += row_bytes) {
          for (int kb = 0; kb < p.nkb; ++kb) tma_load_4d(dst + kb * p.a_kb_stride, &tmA, full, kb * p.swb, 0, h, n);
          if (h == -2) {
            h = -1;  // g = 1: the zero row above image 0
          } else if (++h == p.H) {
            h = -1;  // shared zero row between images
            ++n;
          }
        }
      }
    }
  } else if (warp == 2) {
    // =============================== TMA producer: weights =================================
    if (elect_one()) {
      const int n_w0 = 9 * p.nkb, n_w1 = p.n_chunks * p.nkb1;
      if (p.w0_res || p.w1_res) {
        const uint32_t full = smem_u32(&bar->res_full);
        mbar_expect_tx(full, (p.w0_res ? n_w0 * p.w0_block_bytes : 0) + (p.w1_res ? n_w1 * p.w1_block_bytes : 0));
        if (p.w0_res)
          for (int b = 0; b < n_w0; ++b) tma_load_2d(sbase + p.off_w0 + b * p.w0_block_bytes, &tmW0, full, 0, b * p.OC);
        if (p.w1_res)
          for (int b = 0; b < n_w1; ++b) tma_load_2d(sbase + p.off_w1 + b * p.w1_block_bytes, &tmW1, full, 0, b * p.nc1);
      }
      if (!p.w0_res || !p.w1_res) {
        uint32_t s = 0, ph = 1;  // stage cursor and the parity to wait for on b_empty
        for (int it = 0; it <= n_local; ++it) {  // same interleaving as the MMA thread below
          if (it < n_local && !p.w0_res)
            for (int b = 0; b < n_w0; ++b) {
              mbar_wait(smem_u32(&bar->b_empty[s]), ph);
              mbar_expect_tx(smem_u32(&bar->b_full[s]), p.w0_block_bytes);
              tma_load_2d(sbase + p.off_b + s * p.b_stage_bytes, &tmW0, smem_u32(&bar->b_full[s]), 0, b * p.OC);
              if (++s == (uint32_t)p.SB) { s = 0; ph ^= 1; }
            }
          if (it >= 1 && !p.w1_res)
            for (int b = 0; b < n_w1; ++b) {
              mbar_wait(smem_u32(&bar->b_empty[s]), ph);
              mbar_expect_tx(smem_u32(&bar->b_full[s]), p.w1_block_bytes);
              tma_load_2d(sbase + p.off_b + s * p.b_stage_bytes, &tmW1, smem_u32(&bar->b_full[s]), 0, b * p.nc1);
              if (++s == (uint32_t)p.SB) { s = 0; ph ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t idesc0 = make_idesc_i8(kTileM, p.OC, 0, 1);
      const uint32_t idesc1 = make_idesc_i8(kTileM, p.nc1, 0, 1);
      // descriptors differ only in their 14-bit start-address field (16 B units): constant part + adds
      const uint64_t desc0_hi = make_smem_desc(0, 16, 8 * p.swb, layout_of(p.swb));
      const uint64_t desc1_hi = make_smem_desc(0, 16, 8 * p.swb1, layout_of(p.swb1));
      const uint32_t a_step_kw = p.swb >> 4, a_step_kh = (p.Wp * p.swb) >> 4, a_step_kb = p.a_kb_stride >> 4;
      const uint32_t w0_step = p.w0_block_bytes >> 4, w1_step = p.w1_block_bytes >> 4;
      const uint32_t mid_step_kb = p.mid_kb_stride >> 4, b_stage_step = p.b_stage_bytes >> 4;
      const uint64_t b_stage0 = (sbase + p.off_b) >> 4;
      const int nks_full = p.swb >> 5, nks1_full = p.swb1 >> 5;
      if (p.w0_res || p.w1_res) mbar_wait(smem_u32(&bar->res_full), 0);
      uint32_t bs = 0, bph = 0;  // weight stage cursor / parity to wait for on b_full
      uint32_t c1count = 0;
      Tracer tr(p, 1);
      tr.ev(9);
      for (int it = 0; it <= n_local; ++it) {
        if (it < n_local) {
          // ---- GEMM1(it): acc0 = sum over 9 taps, K-blocks of halo(tile) x W0
          const int tile = blockIdx.x + it * gridDim.x;
          const int sa = it % p.SA;
          const int ab = it % p.n_acc0;
          mbar_wait(smem_u32(&bar->acc0_empty[ab]), ((it / p.n_acc0) & 1) ^ 1);
          mbar_wait(smem_u32(&bar->a_full[sa]), (it / p.SA) & 1);
          tc_fence_after_sync();
          tr.ev(10);
          const int q0 = q_first + tile * kTileM;
          const int g_lo = (q0 - p.Wp - 1) / p.Wp;
          const int a_off_px = (q0 - p.Wp - 1) - g_lo * p.Wp;
          const uint32_t d_tmem = tmem + ab * p.OC;
          uint64_t a_row = desc0_hi | ((sbase + p.off_a + sa * p.a_stage_bytes + a_off_px * p.swb) >> 4);
          uint64_t b_res = desc0_hi | ((sbase + p.off_w0) >> 4);
          uint32_t accumulate = 0;
          for (int kh = 0; kh < 3; ++kh, a_row += a_step_kh) {
            uint64_t a_tap = a_row;
            for (int kw = 0; kw < 3; ++kw, a_tap += a_step_kw) {
              uint64_t a_kb = a_tap;
              for (int kb = 0; kb < p.nkb; ++kb, a_kb += a_step_kb) {
                uint64_t b_desc;
                if (p.w0_res) {
                  b_desc = b_res;
                  b_res += w0_step;
                } else {
                  mbar_wait(smem_u32(&bar->b_full[bs]), bph);
                  tc_fence_after_sync();
                  b_desc = desc0_hi | (b_stage0 + bs * b_stage_step);
                }
                const int nks = (kb == p.nkb - 1) ? p.ks_last : nks_full;
                uint64_t a_ks = a_kb;
                for (int ks = 0; ks < nks; ++ks, a_ks += 2, b_desc += 2) {
                  umma_i8(d_tmem, a_ks, b_desc, idesc0, accumulate);
                  accumulate = 1;
                }
                if (!p.w0_res) {
                  umma_commit(smem_u32(&bar->b_empty[bs]));
                  if (++bs == (uint32_t)p.SB) { bs = 0; bph ^= 1; }
                }
              }
            }
          }
          umma_commit(smem_u32(&bar->a_empty[sa]));
          umma_commit(smem_u32(&bar->acc0_full[ab]));
          tr.ev(11);
        }
        if (it >= 1) {
          // ---- GEMM2(it-1): acc1[chunk] = mid x W1[chunk]
          const int jt = it - 1, mb = jt % p.NM;
          mbar_wait(smem_u32(&bar->mid_full[mb]), (jt / p.NM) & 1);
          tc_fence_after_sync();
          tr.ev(12);
          const uint64_t mid_desc = desc1_hi | ((sbase + p.off_mid + mb * p.mid_bytes) >> 4);
          uint64_t b_res = desc1_hi | ((sbase + p.off_w1) >> 4);
          for (int j = 0; j < p.n_chunks; ++j, ++c1count) {
            const int cb = c1count & 1;
            mbar_wait(smem_u32(&bar->acc1_empty[cb]), ((c1count >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem + kAcc1Col + cb * kAcc1Stride;
            uint64_t a_kb = mid_desc;
            uint32_t accumulate = 0;
            for (int kb = 0; kb < p.nkb1; ++kb, a_kb += mid_step_kb) {
              uint64_t b_desc;
              if (p.w1_res) {
                b_desc = b_res;
                b_res += w1_step;
              } else {
                mbar_wait(smem_u32(&bar->b_full[bs]), bph);
                tc_fence_after_sync();
                b_desc = desc1_hi | (b_stage0 + bs * b_stage_step);
              }
              const int nks = (kb == p.nkb1 - 1) ? p.ks1_last : nks1_full;
              uint64_t a_ks = a_kb;
              for (int ks = 0; ks < nks; ++ks, a_ks += 2, b_desc += 2) {
                umma_i8(d_tmem, a_ks, b_desc, idesc1, accumulate);
                accumulate = 1;
              }
              if (!p.w1_res) {
                umma_commit(smem_u32(&bar->b_empty[bs]));
                if (++bs == (uint32_t)p.SB) { bs = 0; bph ^= 1; }
              }
            }
            umma_commit(smem_u32(&bar->acc1_full[cb]));
            tr.ev(13);
          }
          umma_commit(smem_u32(&bar->mid_empty[mb]));
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ====================================== epilogue =======================================
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - kEpiWarp0) >> 2;  // which half of the 16-column groups
    const int m = quarter * 32 + lane;         // tile row = TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const float* sb0 = reinterpret_cast<const float*>(smem + p.off_bias0);
    const float* ss0 = reinterpret_cast<const float*>(smem + p.off_scale0);
    const float* sb1 = reinterpret_cast<const float*>(smem + p.off_bias1);
    const float* ss1 = reinterpret_cast<const float*>(smem + p.off_scale1);
    const int ts = (kDst == DF_F32 || kDst == DF_S32) ? 4 : 1;
    const uint32_t swz_mask1 = (uint32_t)(p.swb1 / 16 - 1);
    uint32_t c1count = 0;
    Tracer tr(p, 3);
    if (threadIdx.x != kEpiWarp0 * 32) tr.base = nullptr;
    for (int it = 0; it < n_local; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      // where does this row go?
      const int q = q_first + tile * kTileM + m;
      const int g = q / p.Wp, wq = q - g * p.Wp;
      const int n = (g - 1) / p.Hp, hp = (g - 1) - n * p.Hp;
      const bool valid = (wq < p.W) && (hp >= 1) && (n < p.N);
      uint8_t* out_row = static_cast<uint8_t*>(p.dst) + ((size_t)(n * p.H + hp - 1) * p.W + wq) * p.OC1 * ts;

      // ---- conv0 epilogue: acc0 -> u8 intermediate in smem (K-major, swizzled)
      const int ab = it % p.n_acc0, mb = it % p.NM;
      mbar_wait_warp(smem_u32(&bar->mid_empty[mb]), ((it / p.NM) & 1) ^ 1);
      mbar_wait_warp(smem_u32(&bar->acc0_full[ab]), (it / p.n_acc0) & 1);
      tc_fence_after_sync();
      tr.ev(30);
      uint8_t* mid = smem + p.off_mid + mb * p.mid_bytes;
      {
        // two statically indexed register buffers: the TMEM load of group c+2 is in flight while
        // group c is converted (a runtime-indexed acc[buf][] would live in local memory)
        const int nch = p.OC >> 4;
        const uint32_t t_base = lane_addr + ab * p.OC;
        uint32_t acc_a[16], acc_b[16];
        auto emit = [&](const uint32_t* acc, int c) {
          const float4* b4 = reinterpret_cast<const float4*>(sb0 + c * 16);
          const float4* s4 = reinterpret_cast<const float4*>(ss0 + c * 16);
          uint4 v;
          v.x = requant_u8x4<kDown0, kNanSafe>(acc + 0, b4[0], s4[0]);
          v.y = requant_u8x4<kDown0, kNanSafe>(acc + 4, b4[1], s4[1]);
          v.z = requant_u8x4<kDown0, kNanSafe>(acc + 8, b4[2], s4[2]);
          v.w = requant_u8x4<kDown0, kNanSafe>(acc + 12, b4[3], s4[3]);
          const int kb = (c * 16) / p.swb1;
          uint32_t off = (uint32_t)m * p.swb1 + (uint32_t)(c * 16 - kb * p.swb1);
          off ^= ((off >> 7) & swz_mask1) << 4;  // Swizzle<B,4,3> on the (1024 B aligned) block offset
          *reinterpret_cast<uint4*>(mid + kb * p.mid_kb_stride + off) = v;
        };
        if (half < nch) tmem_ld_x16(t_base + half * 16, acc_a);
        for (int c = half; c < nch; c += 4) {
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld_x16(t_base + (c + 2) * 16, acc_b);
          emit(acc_a, c);
          if (c + 2 < nch) {
            tmem_ld_wait();
            if (c + 4 < nch) tmem_ld_x16(t_base + (c + 4) * 16, acc_a);
            emit(acc_b, c + 2);
          }
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar->acc0_empty[ab]));
        mbar_arrive(smem_u32(&bar->mid_full[mb]));
      }
      tr.ev(31);

      // ---- conv1 epilogue: acc1 chunks -> global
      for (int j = 0; j < p.n_chunks; ++j, ++c1count) {
        const int cb = c1count & 1;
        mbar_wait_warp(smem_u32(&bar->acc1_full[cb]), (c1count >> 1) & 1);
        tc_fence_after_sync();
        tr.ev(32);
        int nch = (p.OC1 - j * p.nc1) >> 4;  // real 16-column groups in this chunk
        if (nch > (p.nc1 >> 4)) nch = p.nc1 >> 4;
        const uint32_t t_base = lane_addr + kAcc1Col + cb * kAcc1Stride;
        uint32_t acc_a[16], acc_b[16];
        auto emit = [&](const uint32_t* acc, int c) {
          const int col = j * p.nc1 + c * 16;
          if (valid)
            store16<kDst, kDown1, kNanSafe>(acc, sb1 + col, ss1 + col, p.relu1 != 0, out_row + (size_t)col * ts);
        };
        if (half < nch) tmem_ld_x16(t_base + half * 16, acc_a);
        for (int c = half; c < nch; c += 4) {
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld_x16(t_base + (c + 2) * 16, acc_b);
          emit(acc_a, c);
          if (c + 2 < nch) {
            tmem_ld_wait();
            if (c + 4 < nch) tmem_ld_x16(t_base + (c + 4) * 16, acc_a);
            emit(acc_b, c + 2);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar->acc1_empty[cb]));
        tr.ev(33);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 3) tmem_dealloc<512>(tmem);
}

// ================================================================================ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

CUtensorMapSwizzle swizzle_enum(int swb) {
  return swb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
inline int pick_swb(int k) { return k > 64 ? 128 : (k > 32 ? 64 : 32); }

typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params);

KernelFn pick_kernel(int dst_dt, bool down0, bool down1, bool nan_safe) {
#define DF_PICK(DT)                                                                                     \
  if (dst_dt == DT) {                                                                                   \
    if (nan_safe) {                                                                                     \
      if (down0) return down1 ? conv_fused_kernel<DT, true, true, true> : conv_fused_kernel<DT, true, false, true>; \
      return down1 ? conv_fused_kernel<DT, false, true, true> : conv_fused_kernel<DT, false, false, true>;          \
    }                                                                                                   \
    if (down0) return down1 ? conv_fused_kernel<DT, true, true, false> : conv_fused_kernel<DT, true, false, false>; \
    return down1 ? conv_fused_kernel<DT, false, true, false> : conv_fused_kernel<DT, false, false, false>;          \
  }
  DF_PICK(DF_U8)
  DF_PICK(DF_S8)
  DF_PICK(DF_S32)
  DF_PICK(DF_F32)
#undef DF_PICK
  return nullptr;
}

}  // namespace

struct df_conv {
  df_conv_desc desc;
  Params prm;        // everything except n-dependent fields and dst
  KernelFn kernel;
  uint32_t smem_bytes;
  int device, sms;
  int8_t *d_w0, *d_w1;
  float *d_bias0, *d_scale0, *d_bias1, *d_scale1;
  CUtensorMap tmW0, tmW1;
  // cached activation map (re-encoded when the source pointer or batch changes)
  CUtensorMap tmA;
  const void* tmA_src;
  int tmA_n;
  unsigned long long* trace;
  int trace_cap;
};

namespace {

int is_io_dt(int dt) { return dt == DF_F32 || dt == DF_S32 || dt == DF_S8 || dt == DF_U8; }

// Acceptance rules of the reference: op_conv<T>::init_conf (src/op_conv.cc:262-365) and
// jit_conv_kernel::init_conf (src/jit_conv_kernel.cc:512-673), with defect D1 fixed (oc is the
// 3x3 weight's output channel count; the 1x1 weight is (oc1, oc, 1, 1)).
int validate(const df_conv_desc* d) {
  if (!d) return df::fail(DF_E_INVALID, "conv: null descriptor");
  if (d->n <= 0 || d->ih <= 0 || d->iw <= 0 || d->kh <= 0 || d->kw <= 0 || d->sh <= 0 || d->sw <= 0 || d->ph < 0 ||
      d->pw < 0)
    return df::fail(DF_E_INVALID, "conv: non-positive geometry");
  if (!is_io_dt(d->dst_dt)) return df::fail(DF_E_INVALID, "conv: bad dst dtype %d", d->dst_dt);
  if ((d->bia0_dt != DF_UNDEF && !is_io_dt(d->bia0_dt)) || (d->bia1_dt != DF_UNDEF && !is_io_dt(d->bia1_dt)))
    return df::fail(DF_E_INVALID, "conv: bad bias dtype");
  if (d->ic % 16 || d->oc % 16 || d->ic <= 0 || d->oc <= 0)
    return df::fail(DF_E_INVALID, "conv: ic and oc must be positive multiples of 16 (got %d, %d)", d->ic, d->oc);
  if (d->oc1 < 0 || d->oc1 % 16) return df::fail(DF_E_INVALID, "conv: oc1x1 must be a multiple of 16 (got %d)", d->oc1);
  if ((d->round0 != DF_ROUND_NEAREST && d->round0 != DF_ROUND_DOWN) ||
      (d->round1 != DF_ROUND_NEAREST && d->round1 != DF_ROUND_DOWN))
    return df::fail(DF_E_INVALID, "conv: bad round mode");
  if (d->nscale0 != 1 && d->nscale0 != d->oc) return df::fail(DF_E_INVALID, "conv: conv0 scales must number 1 or oc");
  if (d->oc1 && d->nscale1 != 1 && d->nscale1 != d->oc1)
    return df::fail(DF_E_INVALID, "conv: conv1 scales must number 1 or oc1x1");
  const int oh = (d->ih + 2 * d->ph - d->kh) / d->sh + 1, ow = (d->iw + 2 * d->pw - d->kw) / d->sw + 1;
  if (oh <= 0 || ow <= 0) return df::fail(DF_E_INVALID, "conv: empty output");
  // ur_w based padding limit (jit_conv_kernel.cc:647-661)
  const int nb_oc = d->oc / 16;
  int nb_oc_blocking = nb_oc > 4 ? 4 : nb_oc;
  while (nb_oc % nb_oc_blocking) --nb_oc_blocking;
  int ur_w = 28 / (nb_oc_blocking + 1);
  if (ow < ur_w) ur_w = ow;
  const int tail = ow % ur_w;
  int r_pad_no_tail = (ow - tail - 1) * d->sw + d->kw - d->iw - d->pw;
  if (r_pad_no_tail < 0) r_pad_no_tail = 0;
  if (d->pw > ur_w || r_pad_no_tail > ur_w) return df::fail(DF_E_INVALID, "conv: padding exceeds the register tile");
  return 0;
}

float bias_to_f32(int dt, const void* b, int i) {
  switch (dt) {  // vpmovsxbd / vpmovzxbd / vcvtdq2ps (jit_conv_kernel.cc:238-254): exact or RN
    case DF_F32: return static_cast<const float*>(b)[i];
    case DF_S32: return (float)static_cast<const int32_t*>(b)[i];
    case DF_S8: return (float)static_cast<const int8_t*>(b)[i];
    case DF_U8: return (float)static_cast<const uint8_t*>(b)[i];
    default: return 0.f;
  }
}

// byte offset of (o, i, h, w) in OIhw4i16o4i (jit_conv_kernel.cc:333-338)
size_t blocked_off(int o, int i, int h, int w, int ic, int kh, int kw) {
  const size_t blk = (((size_t)(o / 16) * (ic / 16) + i / 16) * kh + h) * kw + w;
  return blk * 256 + (size_t)((i % 16) / 4) * 64 + (size_t)(o % 16) * 4 + (i % 4);
}

int encode_2d(CUtensorMap* tm, void* base, int row_bytes, long rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return df::fail(DF_E_NODRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gd[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
  cuuint64_t gs[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)row_bytes, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_enum(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return df::fail(DF_E_INTERNAL, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return 0;
}

}  // namespace

extern "C" int df_conv_create(const df_conv_desc* d, const int8_t* wei, const int8_t* wei1, const void* bia0,
                              const void* bia1, const float* scale0, const float* scale1, df_conv** out) {
  if (!out) return df::fail(DF_E_INVALID, "conv: null out");
  *out = nullptr;
  int rc = validate(d);
  if (rc) return rc;
  if (!wei || !scale0) return df::fail(DF_E_INVALID, "conv: null weights / scales");
  if ((d->bia0_dt != DF_UNDEF && !bia0) || (d->bia1_dt != DF_UNDEF && !bia1))
    return df::fail(DF_E_INVALID, "conv: bias dtype given but pointer is null");
  // ---- the B200 path: fused 3x3 s1 p1 + 1x1
  if (d->oc1 == 0) return df::fail(DF_E_UNSUPPORTED, "conv0-only operator is not on the B200 path yet");
  if (!wei1 || !scale1) return df::fail(DF_E_INVALID, "conv: null 1x1 weights / scales");
  if (d->kh != 3 || d->kw != 3 || d->sh != 1 || d->sw != 1 || d->ph != 1 || d->pw != 1)
    return df::fail(DF_E_UNSUPPORTED, "B200 path supports k3 s1 p1 only");
  if (d->oc > 256) return df::fail(DF_E_UNSUPPORTED, "B200 path supports conv0 oc <= 256 (got %d)", d->oc);
  if (d->iw > 254) return df::fail(DF_E_UNSUPPORTED, "B200 path supports width <= 254 (TMA box limit)");

  df_conv* op = new df_conv();
  memset(op, 0, sizeof(*op));
  op->desc = *d;
  Params& p = op->prm;
  p.H = d->ih;
  p.W = d->iw;
  p.IC = d->ic;
  p.OC = d->oc;
  p.OC1 = d->oc1;
  p.swb = pick_swb(d->ic);
  p.nkb = (d->ic + p.swb - 1) / p.swb;
  p.ks_last = (d->ic - (p.nkb - 1) * p.swb + 31) / 32;
  p.swb1 = pick_swb(d->oc);
  p.nkb1 = (d->oc + p.swb1 - 1) / p.swb1;
  p.ks1_last = (d->oc - (p.nkb1 - 1) * p.swb1 + 31) / 32;
  p.Hp = d->ih + 1;
  const int wp_align = 128 / p.swb;  // every halo row must start 128 B aligned for TMA
  p.Wp = (d->iw + 1 + wp_align - 1) / wp_align * wp_align;
  p.NR = (kTileM + 2 * p.Wp) / p.Wp + 2;  // rows touched by 128 + 2*Wp + 2 consecutive positions
  p.nc1 = d->oc1 < 128 ? d->oc1 : 128;
  p.n_chunks = (d->oc1 + p.nc1 - 1) / p.nc1;
  p.n_acc0 = d->oc <= 128 ? 2 : 1;
  p.relu1 = d->relu1;
  p.round0 = d->round0;
  p.round1 = d->round1;

  // ---- shared memory plan
  const int oc1_pad = p.n_chunks * p.nc1;
  uint32_t off = 1024;  // barriers
  p.off_bias0 = off;
  off += align_up(p.OC * 4, 128);
  p.off_scale0 = off;
  off += align_up(p.OC * 4, 128);
  p.off_bias1 = off;
  off += align_up(oc1_pad * 4, 128);
  p.off_scale1 = off;
  off += align_up(oc1_pad * 4, 128);
  off = align_up(off, 1024);
  p.mid_kb_stride = kTileM * p.swb1;
  p.mid_bytes = align_up(p.nkb1 * p.mid_kb_stride, 1024);
  p.off_mid = off;
  const uint32_t fixed_one_mid = off + p.mid_bytes;
  off += 2 * p.mid_bytes;
  p.NM = 2;
  p.a_kb_stride = (uint32_t)p.NR * p.Wp * p.swb;
  p.a_stage_bytes = align_up(p.nkb * p.a_kb_stride, 1024);
  p.w0_block_bytes = (uint32_t)p.OC * p.swb;
  p.w1_block_bytes = (uint32_t)p.nc1 * p.swb1;
  const uint32_t w0_bytes = align_up(9 * p.nkb * p.w0_block_bytes, 1024);
  const uint32_t w1_bytes = align_up(p.n_chunks * p.nkb1 * p.w1_block_bytes, 1024);
  const uint32_t avail = kSmemLimit - 1024;  // base alignment slack
  uint32_t fixed = off;
  {
    // large shapes: give the second intermediate buffer up before giving weight stages up
    const uint32_t stage = align_up(p.w0_block_bytes > p.w1_block_bytes ? p.w0_block_bytes : p.w1_block_bytes, 1024);
    if (fixed + 2 * p.a_stage_bytes + 3 * stage > avail) {
      p.NM = 1;
      fixed = fixed_one_mid;
    }
  }
  if (fixed + 2 * p.a_stage_bytes + w0_bytes + w1_bytes <= avail) {
    p.w0_res = p.w1_res = 1;
    p.off_w0 = fixed;
    p.off_w1 = fixed + w0_bytes;
    p.off_a = fixed + w0_bytes + w1_bytes;
    int sa = (int)((avail - p.off_a) / p.a_stage_bytes);
    p.SA = sa > kMaxAStages ? kMaxAStages : sa;
    p.SB = 1;
    p.b_stage_bytes = 0;
    p.off_b = p.off_a + p.SA * p.a_stage_bytes;
  } else {
    const uint32_t stage_w0 = align_up(p.w0_block_bytes, 1024);
    const uint32_t stage_both = align_up(p.w0_block_bytes > p.w1_block_bytes ? p.w0_block_bytes : p.w1_block_bytes, 1024);
    p.SA = 2;
    if (fixed + 2 * p.a_stage_bytes + w1_bytes + 3 * stage_w0 <= avail) {
      p.w0_res = 0;
      p.w1_res = 1;
      p.off_w1 = fixed;
      p.off_a = fixed + w1_bytes;
      p.b_stage_bytes = stage_w0;
    } else {
      p.w0_res = p.w1_res = 0;
      p.off_a = fixed;
      p.b_stage_bytes = stage_both;
    }
    p.off_b = p.off_a + p.SA * p.a_stage_bytes;
    if (p.off_b + 2 * p.b_stage_bytes > avail) {
      delete op;
      return df::fail(DF_E_UNSUPPORTED, "conv: shape does not fit the shared-memory plan");
    }
    int sb = (int)((avail - p.off_b) / p.b_stage_bytes);
    p.SB = sb > kMaxBStages ? kMaxBStages : sb;
  }
  op->smem_bytes = p.off_b + p.SB * p.b_stage_bytes + 1024;

  // ---- parameters: weights re-laid out K-major per (tap, K-block); bias -> f32; scales expanded
  std::vector<int8_t> w0((size_t)9 * p.nkb * p.OC * p.swb, 0);
  for (int tap = 0; tap < 9; ++tap)
    for (int o = 0; o < p.OC; ++o)
      for (int i = 0; i < p.IC; ++i) {
        const int kb = i / p.swb;
        w0[(((size_t)tap * p.nkb + kb) * p.OC + o) * p.swb + (i - kb * p.swb)] =
            wei[blocked_off(o, i, tap / 3, tap % 3, p.IC, 3, 3)];
      }
  std::vector<int8_t> w1((size_t)p.n_chunks * p.nkb1 * p.nc1 * p.swb1, 0);
  for (int q = 0; q < p.OC1; ++q)
    for (int o = 0; o < p.OC; ++o) {
      const int j = q / p.nc1, r = q - j * p.nc1, kb = o / p.swb1;
      w1[(((size_t)j * p.nkb1 + kb) * p.nc1 + r) * p.swb1 + (o - kb * p.swb1)] = wei1[blocked_off(q, o, 0, 0, p.OC, 1, 1)];
    }
  std::vector<float> b0(p.OC), s0(p.OC), b1(p.OC1), s1(p.OC1);
  bool finite = true;
  for (int o = 0; o < p.OC; ++o) {
    b0[o] = d->bia0_dt != DF_UNDEF ? bias_to_f32(d->bia0_dt, bia0, o) : 0.f;  // x + (+0.0f) == x here
    s0[o] = scale0[d->nscale0 > 1 ? o : 0];                                    // broadcast (defect D4)
    finite = finite && isfinite(b0[o]) && isfinite(s0[o]);
  }
  for (int q = 0; q < p.OC1; ++q) {
    b1[q] = d->bia1_dt != DF_UNDEF ? bias_to_f32(d->bia1_dt, bia1, q) : 0.f;
    s1[q] = scale1[d->nscale1 > 1 ? q : 0];
    finite = finite && isfinite(b1[q]) && isfinite(s1[q]);
  }
  p.nan_safe = !finite;

#define DF_TRY(expr)                          \
  do {                                        \
    int rc_ = (expr);                         \
    if (rc_) {                                \
      df_conv_destroy(op);                    \
      return rc_;                             \
    }                                         \
  } while (0)
#define DF_TRY_CUDA(expr)                                                                        \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      df_conv_destroy(op);                                                                       \
      return df::fail((int)e_, "%s failed: %s", #expr, cudaGetErrorString(e_));                  \
    }                                                                                            \
  } while (0)

  DF_TRY_CUDA(cudaGetDevice(&op->device));
  DF_TRY_CUDA(cudaDeviceGetAttribute(&op->sms, cudaDevAttrMultiProcessorCount, op->device));
  DF_TRY_CUDA(cudaMalloc(&op->d_w0, w0.size()));
  DF_TRY_CUDA(cudaMalloc(&op->d_w1, w1.size()));
  DF_TRY_CUDA(cudaMalloc(&op->d_bias0, p.OC * 4));
  DF_TRY_CUDA(cudaMalloc(&op->d_scale0, p.OC * 4));
  DF_TRY_CUDA(cudaMalloc(&op->d_bias1, p.OC1 * 4));
  DF_TRY_CUDA(cudaMalloc(&op->d_scale1, p.OC1 * 4));
  DF_TRY_CUDA(cudaMemcpy(op->d_w0, w0.data(), w0.size(), cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_w1, w1.data(), w1.size(), cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_bias0, b0.data(), p.OC * 4, cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_scale0, s0.data(), p.OC * 4, cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_bias1, b1.data(), p.OC1 * 4, cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_scale1, s1.data(), p.OC1 * 4, cudaMemcpyHostToDevice));
  p.bias0 = op->d_bias0;
  p.scale0 = op->d_scale0;
  p.bias1 = op->d_bias1;
  p.scale1 = op->d_scale1;
  DF_TRY(encode_2d(&op->tmW0, op->d_w0, p.swb, (long)9 * p.nkb * p.OC, p.OC));
  DF_TRY(encode_2d(&op->tmW1, op->d_w1, p.swb1, (long)p.n_chunks * p.nkb1 * p.nc1, p.nc1));

  op->kernel = pick_kernel(d->dst_dt, d->round0 == DF_ROUND_DOWN, d->round1 == DF_ROUND_DOWN, p.nan_safe != 0);
  DF_TRY_CUDA(cudaFuncSetAttribute((const void*)op->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op->smem_bytes));
  *out = op;
  return 0;
}

static int tiles_for(const Params& p, int n) {
  const long q_first = 2L * p.Wp, q_end = ((long)n * p.Hp + 1) * p.Wp;
  return (int)((q_end - q_first + kTileM - 1) / kTileM);
}

extern "C" int df_conv_run(df_conv* op, const uint8_t* src, void* dst, int n, void* stream) {
  if (!op || !src || !dst) return df::fail(DF_E_INVALID, "conv run: null argument");
  if (n < 0 || n > op->desc.n) return df::fail(DF_E_INVALID, "conv run: batch %d outside [0, %d]", n, op->desc.n);
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return df::fail(DF_E_INVALID, "conv run: src/dst must be 16-byte aligned");
  Params p = op->prm;
  if ((long)n * p.Hp * p.Wp + 4L * p.Wp + kTileM >= (1L << 31))
    return df::fail(DF_E_UNSUPPORTED, "conv run: batch too large for 32-bit position index");
  if (op->tmA_src != src || op->tmA_n != n) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return df::fail(DF_E_NODRIVER, "cuTensorMapEncodeTiled unavailable");
    cuuint64_t gd[4] = {(cuuint64_t)p.IC, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)n};
    cuuint64_t gs[3] = {(cuuint64_t)p.IC, (cuuint64_t)p.W * p.IC, (cuuint64_t)p.H * p.W * p.IC};
    cuuint32_t box[4] = {(cuuint32_t)p.swb, (cuuint32_t)p.Wp, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&op->tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(src), gd, gs, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(p.swb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return df::fail(DF_E_INTERNAL, "cuTensorMapEncodeTiled(src) failed: %d", (int)r);
    op->tmA_src = src;
    op->tmA_n = n;
  }
  p.N = n;
  p.n_tiles = tiles_for(p, n);
  p.dst = dst;
  p.trace = op->trace;
  p.trace_cap = op->trace_cap;
  const int grid = p.n_tiles < op->sms ? p.n_tiles : op->sms;
  op->kernel<<<grid, kThreads, op->smem_bytes, (cudaStream_t)stream>>>(op->tmA, op->tmW0, op->tmW1, p);
  DF_CUDA(cudaGetLastError());
  return 0;
}

// Diagnostic: record a per-role clock64 timeline into `dev_buf` (grid * 4 * cap u64 words) on the
// following launches; pass null to switch it off.  Not part of the reference-facing surface.
extern "C" int df_conv_debug_trace(df_conv* op, void* dev_buf, int cap) {
  if (!op) return df::fail(DF_E_INVALID, "trace: null op");
  op->trace = static_cast<unsigned long long*>(dev_buf);
  op->trace_cap = dev_buf ? cap : 0;
  return 0;
}

extern "C" int df_conv_query(const df_conv* op, df_conv_info* info) {
  if (!op || !info) return df::fail(DF_E_INVALID, "conv query: null argument");
  const Params& p = op->prm;
  info->tiles_per_launch = tiles_for(p, op->desc.n);
  info->grid = info->tiles_per_launch < op->sms ? info->tiles_per_launch : op->sms;
  info->block = kThreads;
  info->smem_bytes = (int)op->smem_bytes;
  info->w0_resident = p.w0_res;
  info->w1_resident = p.w1_res;
  info->a_stages = p.SA;
  info->b_stages = p.SB;
  info->padded_w = p.Wp;
  info->padded_h = p.Hp;
  info->macs_per_image = (double)p.H * p.W * (9.0 * p.IC * p.OC + (double)p.OC * p.OC1);
  info->mma_efficiency = (double)op->desc.n * p.H * p.W / ((double)info->tiles_per_launch * kTileM);
  return 0;
}

extern "C" int df_conv_destroy(df_conv* op) {
  if (!op) return 0;
  cudaFree(op->d_w0);
  cudaFree(op->d_w1);
  cudaFree(op->d_bias0);
  cudaFree(op->d_scale0);
  cudaFree(op->d_bias1);
  cudaFree(op->d_scale1);
  delete op;
  return 0;
}
