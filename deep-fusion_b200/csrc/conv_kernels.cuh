// conv_kernels.cuh -- device code of conv_fused.cu: fused u8 x s8 conv3x3(s1,p1)+ReLU+conv1x1+ReLU for sm_100a (B200).
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
//  * conv0 accumulates in TMEM; the 16 epilogue warps apply (float(acc)+bias)*scale -> ReLU -> round
//    -> u8 and write the tile straight into shared memory in the swizzled K-major layout the
//    second GEMM wants, so the intermediate never leaves the SM.
//  * conv1x1 runs as N-chunks of <=128 output channels through two TMEM accumulators, so the
//    epilogue of chunk j overlaps the MMA of chunk j+1 and the conv0 MMAs of the next tile.
//  * the epilogue reads row-pair TMEM fragments of channel-permuted accumulators (one set of
//    per-channel constants serves four rows), stages 1-byte output in shared memory and sends it
//    out with two TMA stores per chunk (see DstMaps / store_staged_chunk).
//  * warp roles: w0 TMA(A) | w1 GEMM1 issue | w2 TMA(weights), then GEMM2 issue | w3 TMEM alloc,
//    then TMA stores | w4-19 epilogue.  All hand-offs are mbarriers; tcgen05.commit releases stages.
//  * weights live in shared memory for the whole (persistent) kernel when they fit -- for the
//    cfg3 shape split across a CTA pair (conv_pair_kernel, cta_group::2); otherwise they stream
//    through a ring of stages in exactly the order the MMA thread consumes them.
//  * launches use programmatic dependent launch: the next launch's prologue runs under this
//    launch's tail (griddep_launch_dependents / griddep_wait).
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "df_common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace dfconv {

constexpr int kEpiWarp0 = 4;    // warps 0..3: TMA(A) | MMA | TMA(W) | TMEM alloc
constexpr int kEpiWarps = 16;   // epilogue warps (four per TMEM lane quarter)
// 1 (default): all 16 warps on one unit.  2: two groups of 8 on alternate units -- measured slower on cfg3 /
// cfg4 (19.4 vs 18.2 us, 98 vs 89 us), slightly faster on cfg1 (32.7 vs 35.2 us); kept as a build knob.
#ifndef DF_EPI_GROUPS
#define DF_EPI_GROUPS 1
#endif
constexpr int kEpiGroups = DF_EPI_GROUPS;             // groups of epilogue warps; group u % kEpiGroups runs unit u
constexpr int kUnitWarps = kEpiWarps / kEpiGroups;    // warps that share one work unit (= arrivals per hand-off)
constexpr int kStageBufs = 2;   // output staging buffers (unit c uses buffer c & 1)
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);  // 640: a 21st warp would round the register
                                                        // allocation up to 24 warps (80 registers per thread)
constexpr int kTileM = 128;
constexpr int kMaxAStages = 4;
constexpr int kMaxBStages = 8;
constexpr int kMaxSrc = 8;       // inputs of a fused concat -> conv
constexpr int kMaxKBlocks = 16;  // K-blocks of the conv0 reduction when the input is a fused concat
constexpr int kAcc1Col = 256;   // TMEM column of the first conv1 accumulator
constexpr int kAcc1Stride = 128;
constexpr uint32_t kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct Params {
  int N, H, W, IC, OC, OC1;
  // conv0 window: KH x KW taps, padding PH / PW, stride 1 (2 PH <= KH - 1, 2 PW <= KW - 1, so the output is not
  // larger than the input: OH = H + 2 PH - KH + 1 <= H, likewise OW).  ZR = max(PH, KH - 1 - PH) zero rows separate
  // the images of the linearised padded pixel space (Hp = H + ZR); q_first = (ZR + 1) * Wp is the position of image
  // 0, row 0, column 0.  The BASELINE geometry is 3, 3, 1, 1 -> OH = H, OW = W, ZR = 1, q_first = 2 Wp.
  int KH, KW, PH, PW, OH, OW, ZR, q_first;
  // Strided windows (SH, SW > 1; run-time geometry only) are computed as the stride-1 convolution and only every
  // SH-th row / SW-th column of it is stored: OHS x OWS = the operator's real output size.  (Correct and simple,
  // at SH * SW times the arithmetic; the strided convs of the networks this serves are few and small.)
  int SH, SW, OHS, OWS;
  // destination row pitch in channels and first channel: an operator with more output channels than one TMEM
  // accumulator holds (256) runs as several launches, each writing its own channel range of the same pixels
  int dst_pitch, dst_ch0;
  // halo rows wider than the TMA box limit (256 positions) are loaded as n_box boxes of box_w positions each
  int n_box, box_w;
  int Hp, Wp, NR;
  int n_tiles;
  int swb, nkb, ks_last;     // conv0: K-block bytes (= swizzle span), blocks, 32 B steps in last
  int swb1, nkb1, ks1_last;  // conv1
  int nc1, n_chunks, n_acc0;
  int SA, SB, NM, w0_res, w1_res;  // halo stages, weight stages, intermediate buffers
  int tile_step_mod;               // (128 * gridDim.x) mod Wp: halo-window offset step per tile
  // position steps for the epilogue's division-free bookkeeping (PosState): from a tile to the next tile of
  // the same CTA, and 8 positions down inside a tile; each as (columns, images, rows) with
  // step = ((images * Hp + rows) * Wp + columns) positions
  int ts_dw, ts_dn, ts_dh, r8_dw, r8_dn, r8_dh;
  uint32_t off_bias0, off_scale0, off_bias1, off_scale1, off_k1;
  int fast1;      // conv1 int->float by exact offset-magic conversion (k1 / bias1 hold K[q] / C[q])
  int k1_uniform; // non-zero: one K serves every channel (global lower bound keeps the range < 2^23)
  uint32_t off_a, a_stage_bytes, a_kb_stride;
  uint32_t off_mid, mid_bytes, mid_kb_stride;
  uint32_t off_w0, w0_block_bytes, off_w1, w1_block_bytes;
  uint32_t off_b, b_stage_bytes;
  int g_interleave;  // streamed CTA-pair kernel: GEMM2(it - 1)'s chunks and GEMM1(it)'s taps alternate in the ring's order
  int w1_pack;  // streamed CTA-pair kernel: conv1 K-blocks (weight halves) per ring stage (1, or 2 when two fit and nkb1 is even)
  int relu1, round0, round1, nan_safe;
  int dbg_no_mma;      // diagnostic (DF_DEBUG_NO_MMA=1): skip every tcgen05.mma, keep the hand-offs; results are garbage
  int conv0_only;      // conv() without the 1x1 stage (include/deepfusion.h:121-129): the 3x3 accumulator goes
                       // through the conv1 finish (bias1 / scale1 / relu1 / round1 hold the conv0 values,
                       // OC1 == OC, chunks of 128 accumulator columns); run-time geometry only
  int stage_out;       // 1-byte destinations: conv1 chunks are staged in smem and leave by TMA store
  uint32_t off_stage;  // kStageBufs staging buffers of kTileM x 128 B (128 B-swizzled rows)
  const float *bias0, *scale0, *bias1, *scale1;
  const int* k1;
  void* dst;
  // eltwise-sum (+ReLU) fused into the final stage (README.md:65 "eltwise-sum + relu fused op"; run-time geometry
  // only): `res` has the destination's type and layout; t = (float(acc) + bias) * scale, then t = t + float(res)
  // as one more separately rounded f32 add, then ReLU / round / saturate as always.  Null: no sum.
  const void* res;
  unsigned long long* trace;  // optional timeline buffer (df_conv_debug_trace), normally null
  int trace_cap;
  // concat fused into the A-operand load (df_conv_create_concat; run-time geometry only): the conv's input is the
  // channel concatenation of n_src NHWC tensors that are never materialised -- K-block kb of the halo is loaded
  // from tensor map kb_src[kb] at channel kb_c0[kb] (every source's channel count is a multiple of swb, so a
  // K-block never straddles two sources).  concat_relu: the reference's literal u8 ReLU (vpmaxsb: bytes >= 128
  // become 0, jit_concat_kernel.cc:43-51) is applied to the halo in shared memory before the tensor pipe reads it.
  int n_src, concat_relu;
  // A-operand (halo) K-blocks of swa bytes, nka of them: equal to swb / nkb except for a fused concat, whose halo
  // K-blocks are as narrow as the inputs' channel counts require (32 / 64 / 128 B) while the WEIGHTS keep the widest
  // K-blocks -- the two operand descriptors of a tcgen05.mma carry their own swizzle mode, and weight blocks of 4 KB
  // instead of 16 KB would quarter the bytes in flight of the weight ring
  int swa, nka;
  // K-slicing of the halo (single-CTA run-time-geometry kernel): an input too deep for a halo stage (e.g. 7x7 x 2048
  // channels: 286 KB) goes through the stages in n_ks slices of nkb_s K-blocks; GEMM1 accumulates over the slices of
  // a tile in TMEM, weights stream in that order.  n_ks == 1 everywhere else.
  int n_ks, nkb_s;
  unsigned char kb_src[kMaxKBlocks];
  unsigned short kb_c0[kMaxKBlocks];
};

// Diagnostic switches that produce garbage results (skip the MMAs / the TMA stores / the staging writes to
// time what is left) exist only in builds with -DDF_DIAG=1; in production builds they fold to nothing.
#ifndef DF_DIAG
#define DF_DIAG 0
#endif
__device__ __forceinline__ bool dbg_flag(const Params& p, int bit) {
#if DF_DIAG
  return (p.dbg_no_mma & bit) != 0;
#else
  (void)p; (void)bit;
  return false;
#endif
}

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
// The epilogue warps are bound by instruction issue, and even a disabled trace point costs them ~10
// instructions (ncu: 15 % of the epilogue's instructions with five points per unit).  Their trace points
// exist only in builds with -DDF_EPI_TRACE=1 (make DF_NVFLAGS=-DDF_EPI_TRACE=1; scripts/trace_conv.py).
#ifndef DF_EPI_TRACE
#define DF_EPI_TRACE 0
#endif
struct EpiTracer {
#if DF_EPI_TRACE
  Tracer t;
  __device__ EpiTracer(const Params& p, int role, bool on) : t(p, role) {
    if (!on) t.base = nullptr;
  }
  __device__ __forceinline__ void ev(unsigned tag) { t.ev(tag); }
#else
  __device__ EpiTracer(const Params&, int, bool) {}
  __device__ __forceinline__ void ev(unsigned) {}
#endif
};

// kernel entry / exit wall-clock stamps (%globaltimer, ns; comparable across SMs) in the last two words of
// the CTA's role-2 trace lane
__device__ __forceinline__ void trace_wallclock(const Params& p, int slot) {
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[((size_t)blockIdx.x * 4 + 2) * p.trace_cap + p.trace_cap - 1 - slot] = t;
  }
}

// GEMM1 issue throttle.  The tensor pipe executes MMAs in the order they were issued, whichever thread issued
// them.  GEMM1 of the next tile (36 MMAs, 2304 cycles for cfg3) can be issued long before the conv1 chunks of the
// current tile are (they wait for the conv0 epilogue), and a chunk issued behind it is not computed until all of
// it has run -- the epilogue then sits idle for a whole GEMM1 and the two phases serialise (profiles/
// r02_trace_cfg3_seed.log: acc1 chunks ready 1500 cycles after their issue).  So the GEMM1 thread keeps at most
// kG1Ahead tap groups (one tap = K-blocks x K-steps MMAs, 256 cycles for cfg3) in the queue beyond the completed
// ones: after each tap it commits to g1_prog[tap % kG1Ahead] and before issuing tap i it waits for tap
// i - kG1Ahead.  A chunk that becomes ready is then at most kG1Ahead taps away from the tensor pipe.
#ifndef DF_G1_AHEAD
#define DF_G1_AHEAD 2
#endif
constexpr int kG1Ahead = DF_G1_AHEAD;
#ifndef DF_WIDE_ST
#define DF_WIDE_ST 1  // A/B knob: 256-bit stores for 4-byte destinations in the static epilogue
#endif
#ifndef DF_G1_FREE_FIRST
#define DF_G1_FREE_FIRST 1  // A/B knob: 1 = the first tile's GEMM1 is not throttled (CTA-pair kernel)
#endif

// Build knobs for A/B measurements (defaults = the production configuration):
//   DF_SEED        1: the static kernels keep the conv1 accumulators pre-seeded with the offset-magic constant
//   DF_STATIC_EPI  1: the static kernels run epilogue_static (two groups, per-warp TMA stores); 0: epilogue_role
#ifndef DF_SEED
#define DF_SEED 1
#endif
#ifndef DF_STATIC_EPI
#define DF_STATIC_EPI 1
#endif

template <class G>
constexpr bool seeded_acc1() { return G::is_static && DF_SEED != 0; }
template <class G>
constexpr bool static_epilogue() { return G::is_static && DF_STATIC_EPI != 0; }
// DF_STATIC_GROUPS 2: the static epilogue runs as TWO independent groups of eight warps (epilogue_static2): group g
// owns conv1 accumulator g and the chunks that land in it, the conv0 epilogues alternate between the groups.  The
// groups drift apart, so one group's barrier / TMEM-load / re-seed latency chain runs under the other's arithmetic
// instead of all 16 warps walking every unit in lock-step.  1: one group of 16 (epilogue_static).
// 0 (default): two groups only where they measured faster -- conv0 accumulators of <= 64 columns (BASELINE cfg1:
// +1.3 .. 2.9 %; cfg3 -3.5 %, cfg4 -4 .. 7 %, profiles/r02_variants_two_group_epilogue.log).
#ifndef DF_STATIC_GROUPS
#define DF_STATIC_GROUPS 0
#endif
template <class G>
constexpr bool static_two_groups() {
  if constexpr (G::is_static)
    return static_epilogue<G>() && G::n_chunks % 2 == 0 && (DF_STATIC_GROUPS == 2 || (DF_STATIC_GROUPS == 0 && G::OC <= 64));
  else
    return false;
}
// epilogue warps that arrive on a unit's hand-off barriers
template <class G>
constexpr int unit_warps() { return static_two_groups<G>() ? kEpiWarps / 2 : kUnitWarps; }
// DF_SEED_CP 1: the seed is written by the tensor pipe itself -- the GEMM2 issuer puts 32 tcgen05.cp (smem -> TMEM,
// 32 x 128 bit broadcast to the four lane quarters, source = the K vector in shared memory) in front of a chunk's
// MMAs; tcgen05.cp and tcgen05.mma execute in issue order, so no wait is needed and the epilogue's tcgen05.st +
// tcgen05.wait::st (150..250 cycles of every conv1 unit's latency chain) disappear.  0: the epilogue warps re-seed.
#ifndef DF_SEED_CP
#define DF_SEED_CP 0
#endif
template <class G>
constexpr bool seed_by_cp() { return seeded_acc1<G>() && DF_SEED_CP != 0; }
template <class G>
constexpr bool epilogue_seeds() { return seeded_acc1<G>() && DF_SEED_CP == 0; }  // then it also hands the buffers over first
template <bool kPair>
__device__ __forceinline__ void seed_chunk_cp(uint32_t d_tmem, uint64_t k_desc) {
#pragma unroll
  for (int i = 0; i < kAcc1Stride / 4; ++i) {
    if constexpr (kPair)
      asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(d_tmem + 4 * i), "l"(k_desc) : "memory");
    else
      asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(d_tmem + 4 * i), "l"(k_desc) : "memory");
  }
}

struct Barriers {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t a_ready[kMaxAStages];  // fused concat + ReLU: halo stage clamped and visible to the tensor pipe
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];
  uint64_t res_full;
  uint64_t acc0_full[2], acc0_empty[2];
  uint64_t mid_full[2], mid_empty[2];
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t stage_full[2], stage_empty[2];  // staged output: epilogue group <-> store thread (warp 3)
  uint64_t g1_prog[kG1Ahead];              // GEMM1 issue throttle (see G1Throttle)
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t layout_of(int swb) {
  return swb == 128 ? kLayoutSW128 : (swb == 64 ? kLayoutSW64 : kLayoutSW32);
}

constexpr uint32_t kStageBytes = kTileM * 128;  // one staged conv1 chunk: 128 positions x 128 channels (1 byte each)

// Staged output of the conv1 chunks (1-byte destinations).  Plain per-thread stores are the wrong tool
// here: a thread owns one position (row) of the tile, so a warp-wide 16-byte store touches 32 different
// 128-byte lines and the LSU spends ~32 cycles on it (profiles/r01_epilogue_unit_bench.log: stores tripled
// the epilogue time).  Instead the epilogue group writes the chunk into shared memory and the TMA unit
// sends it out:
//   * the destination is viewed as a 2-D tensor [N*H*W pixels][OC1 channels]; the valid positions of a
//     tile are a CONTIGUOUS pixel range [f0, f0 + V) of it (padding positions simply do not exist there);
//   * the thread of valid position q writes its 16-byte units to staging row  valid_before(q) - f0  (rows of
//     128 B, units XOR-swizzled with row & 7 = SWIZZLE_128B), threads of padding positions write nothing;
//   * one thread (warp 3) stores rows [0, P) and [V - P, V), P = largest power of two <= V, with the
//     tensor map whose box is {128 channels, P pixels}: at most two TMA stores per chunk, overlapping rows
//     carry identical bytes.  (Negative start coordinates, which would allow a single clipped box, are
//     rejected by the hardware for stores -- probe/tma_store_probe.cu.)
struct DstMaps {
  CUtensorMap m[8];  // m[i]: box = {128, 128 >> i}
};
// Activation maps: m[0] is the conv's source; m[1 ..] only exist for a fused concat -> conv (Params::n_src)
struct SrcMaps {
  CUtensorMap m[kMaxSrc];
};

struct PosState {
  int wq, n, hp;
};
__device__ __forceinline__ int pos_valid_before(const Params& p, const PosState& s);
// number of valid (= real pixel) positions with linear index < q; for a valid q this is its flat NHW
// pixel index
__device__ __forceinline__ int valid_before(const Params& p, int q) {
  const int gq = q / p.Wp, wq = q - gq * p.Wp;
  const int t = gq - 1;
  const int n = t / p.Hp, hp = t - n * p.Hp;
  PosState s;
  s.wq = wq;
  s.n = n;
  s.hp = hp;
  return pos_valid_before(p, s);
}

// Position q = ((n * Hp + hp) + 1) * Wp + wq of the linearised padded pixel space, kept as (wq, n, hp) and
// advanced by precomputed steps: the epilogue needs the coordinates of four positions per thread and
// tile, and computing them with integer divisions cost more than a conv1 chunk's arithmetic
// (~1500 cycles per tile with 16 warps, profiles/r01_trace_cfg3_v10.log).
__device__ __forceinline__ PosState pos_of(const Params& p, int q) {
  const int gq = q / p.Wp;
  PosState s;
  s.wq = q - gq * p.Wp;
  s.n = (gq - 1) / p.Hp;
  s.hp = (gq - 1) - s.n * p.Hp;
  return s;
}
__device__ __forceinline__ void pos_step(const Params& p, PosState& s, int dw, int dn, int dh) {
  s.wq += dw;
  int carry = 0;
  if (s.wq >= p.Wp) {
    s.wq -= p.Wp;
    carry = 1;
  }
  s.hp += dh + carry;
  s.n += dn;
  if (s.hp >= p.Hp) {
    s.hp -= p.Hp;
    ++s.n;
  }
}
// NHW pixel index of a position, -1 for padding positions
__device__ __forceinline__ int pos_pixel(const Params& p, const PosState& s) {
  const bool ok = (s.wq < p.OW) && (s.hp >= p.ZR) && (s.hp < p.ZR + p.OH) && (s.n < p.N);
  if (p.SH * p.SW > 1) {  // strided window: only every SH-th row / SW-th column is an output pixel
    const int r = s.hp - p.ZR, y = r / p.SH, x = s.wq / p.SW;
    return (ok && y * p.SH == r && x * p.SW == s.wq) ? (s.n * p.OHS + y) * p.OWS + x : -1;
  }
  return ok ? (s.n * p.OH + s.hp - p.ZR) * p.OW + s.wq : -1;
}
// = valid_before(q) for the position's q
__device__ __forceinline__ int pos_valid_before(const Params& p, const PosState& s) {
  if (p.SH * p.SW > 1) {
    if (s.n >= p.N) return p.N * p.OHS * p.OWS;
    const int r = s.hp - p.ZR;
    if (r < 0) return s.n * p.OHS * p.OWS;
    const int rows = min((r + p.SH - 1) / p.SH, p.OHS);  // output rows entirely before this position
    const int y = r / p.SH;
    const bool row_valid = y * p.SH == r && y < p.OHS;
    return (s.n * p.OHS + rows) * p.OWS + (row_valid ? min((s.wq + p.SW - 1) / p.SW, p.OWS) : 0);
  }
  if (s.n >= p.N) return p.N * p.OH * p.OW;
  if (s.hp < p.ZR) return s.n * p.OH * p.OW;
  if (s.hp >= p.ZR + p.OH) return (s.n + 1) * p.OH * p.OW;
  return (s.n * p.OH + s.hp - p.ZR) * p.OW + min(s.wq, p.OW);
}

__device__ __forceinline__ void store_staged_chunk(const DstMaps& dm, uint32_t stage, int f0, int V, int ch0) {
  if (V > 0) {
    const int lg = 31 - __clz(V), P = 1 << lg;
    const CUtensorMap* tm = &dm.m[7 - lg];
    tma_store_2d(tm, stage, ch0, f0);
    if (V != P) tma_store_2d(tm, stage + (uint32_t)(V - P) * 128u, ch0, f0 + V - P);
  }
  bulk_commit_group();
}

// ------------------------------------------------------------------------ geometry policies
// The BASELINE.json shapes get their channel geometry at compile time: the MMA issue loop then
// unrolls completely (9 taps x K-blocks x K-steps), weight-stage indices become constants and all
// descriptor arithmetic is immediate adds in the uniform datapath.  Every other accepted shape runs
// the same code through the run-time policy (loops stay loops; slower issue, same results).
template <int kIC, int kOC, int kOC1, int kW0Res, int kW1Res, int kSB>
struct StaticGeom {
  static constexpr bool is_static = true;
  static constexpr int IC = kIC, OC = kOC, OC1 = kOC1, w0_res = kW0Res, w1_res = kW1Res, SB = kSB;
  static constexpr int swb = kIC > 64 ? 128 : (kIC > 32 ? 64 : 32);
  static constexpr int nkb = (kIC + swb - 1) / swb;
  static constexpr int ks_last = (kIC - (nkb - 1) * swb + 31) / 32;
  static constexpr int swb1 = kOC > 64 ? 128 : (kOC > 32 ? 64 : 32);
  static constexpr int nkb1 = (kOC + swb1 - 1) / swb1;
  static constexpr int ks1_last = (kOC - (nkb1 - 1) * swb1 + 31) / 32;
  static constexpr int nc1 = kOC1 < 128 ? kOC1 : 128;
  static constexpr int n_chunks = (kOC1 + nc1 - 1) / nc1;
  static constexpr int n_acc0 = kOC <= 128 ? 2 : 1;
  static constexpr int KH = 3, KW = 3, PH = 1, PW = 1;
  static constexpr int swa = swb, nka = nkb;
};
struct DynGeom {
  static constexpr bool is_static = false;
};

#define DF_GEO(name)                                  \
  __device__ __forceinline__ int name() const {       \
    if constexpr (G::is_static) return G::name;       \
    else return p.name;                               \
  }
template <class G>
struct Geo {
  const Params& p;
  DF_GEO(IC) DF_GEO(OC) DF_GEO(OC1) DF_GEO(w0_res) DF_GEO(w1_res) DF_GEO(SB)
  DF_GEO(swb) DF_GEO(nkb) DF_GEO(ks_last) DF_GEO(swb1) DF_GEO(nkb1) DF_GEO(ks1_last)
  DF_GEO(nc1) DF_GEO(n_chunks) DF_GEO(n_acc0) DF_GEO(KH) DF_GEO(KW) DF_GEO(PH) DF_GEO(PW) DF_GEO(swa) DF_GEO(nka)
  __device__ __forceinline__ uint32_t w0_block_bytes() const { return (uint32_t)(OC() * swb()); }
  __device__ __forceinline__ uint32_t w1_block_bytes() const { return (uint32_t)(nc1() * swb1()); }
  __device__ __forceinline__ uint32_t mid_kb_stride() const { return (uint32_t)(kTileM * swb1()); }
};
#undef DF_GEO

// ---------------------------------------------------------------------------- epilogue math
// t = (float(acc) + bias) * scale as separately rounded f32 operations -- vcvtdq2ps, vaddps, vmulps
// (jit_conv_kernel.cc:96-100, :259-263), two lanes at a time on the packed f32x2 pipe (add.rn /
// mul.rn keep IEEE rounding per lane; verified bit-exact in probe/umma_probe.cu).  Never an FMA.
__device__ __forceinline__ void scale_pair(uint32_t a0, uint32_t a1, float b0, float b1, float s0, float s1, float& t0,
                                           float& t1) {
  const float f0 = __int2float_rn((int)a0), f1 = __int2float_rn((int)a1);
  unsigned long long f, b, s;
  asm("mov.b64 %0, {%1, %2};" : "=l"(f) : "f"(f0), "f"(f1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(s) : "f"(s0), "f"(s1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(f), "l"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(f), "l"(s));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(f));
}
__device__ __forceinline__ void scale4(const uint32_t* acc, const float4 b, const float4 s, float* t) {
  scale_pair(acc[0], acc[1], b.x, b.y, s.x, s.y, t[0], t[1]);
  scale_pair(acc[2], acc[3], b.z, b.w, s.z, s.w, t[2], t[3]);
}
// conv1 fast path.  acc1 = sum of u8 * s8 over OC <= 256 terms lies in [lo[q], lo[q] + 2^23) with
// lo[q] = 255 * (sum of the negative weights of channel q), so with K[q] = 0x4B000000 - lo[q]
//   __int_as_float(acc + K[q]) == 2^23 + (acc - lo[q])      exactly (ulp is 1 in [2^23, 2^24)),
// and adding C[q] = bias[q] + lo[q] - 2^23 (exactly representable, checked at create time) rounds
// the real number acc + bias ONCE -- bit-identical to vcvtdq2ps ; vaddps, because float(acc) is
// exact below 2^24.  This moves the conversion off the quarter-rate I2F pipe (probe/epi_pipes.cu).
__device__ __forceinline__ void scale4_fast(const uint32_t* acc, const int4 k, const float4 c, const float4 s, float* t) {
  const float f0 = __int_as_float((int)acc[0] + k.x), f1 = __int_as_float((int)acc[1] + k.y);
  const float f2 = __int_as_float((int)acc[2] + k.z), f3 = __int_as_float((int)acc[3] + k.w);
  unsigned long long a, b, cc, dd, s0, s1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f0), "f"(f1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(f2), "f"(f3));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(c.z), "f"(c.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(s0) : "f"(s.x), "f"(s.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(s1) : "f"(s.z), "f"(s.w));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(cc));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(b), "l"(dd));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(s0));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(b), "l"(s1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t[0]), "=f"(t[1]) : "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t[2]), "=f"(t[3]) : "l"(b));
}
__device__ __forceinline__ float4 load_scale4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 lds128f(uint32_t addr) {  // 16-byte shared-memory load from a 32-bit shared address
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
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
__device__ __forceinline__ uint32_t pack_u8x4(const float* t) {
  int q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    q[i] = kDown ? __float2int_rd(t[i]) : __float2int_rn(t[i]);
    if (kNanSafe && t[i] != t[i]) q[i] = 255;
  }
  uint32_t hi, lo;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q[3]), "r"(q[2]));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(q[1]), "r"(q[0]), "r"(hi));
  return lo;
}
template <bool kDown, bool kNanSafe>
__device__ __forceinline__ uint32_t requant_u8x4(const uint32_t* acc, const float4 b, const float4 s) {
  float t[4];
  scale4(acc, b, s, t);
  return pack_u8x4<kDown, kNanSafe>(t);
}

template <bool kDown>
__device__ __forceinline__ uint32_t pack_s8x4(float* t, bool relu) {
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

// ---------------------------------------------------------------- TMEM fragments and channel order
// The epilogue reads accumulators as 16-lane "row pair" fragments: tcgen05.ld.16x256b.xN hands thread t of
// the warp rows (t/4) and (t/4)+8 of a 16-lane block and, for k = 0..N-1, columns 8k + 2(t%4) + {0,1}:
//   r[4k + 2*hl + e] = (row t/4 + 8*hl, column 8k + 2(t%4) + e).
// df_conv_create orders the output channels (= rows of the weight matrices = accumulator columns) such
// that inside a block of NB columns (NB = 32 read with .x4, or NB = 16 read with .x2 for channel counts
// that are not multiples of 32) column 8k + 2m + e holds channel (NB/4)*m + 2k + e.  Thread t therefore
// owns NB/4 CONSECUTIVE channels (starting at (NB/4)*(t%4)) of four rows (two 16-lane loads):
//   * one set of per-channel constants (bias, scale, K) serves four rows -- with one row per thread
//     (32x32b fragments) every thread needs the constants of every column, and feeding them (broadcast
//     LDS: ~1.8 constants/clk/SM, profiles/r01_lds_broadcast_vs_constbank.log) costs more than the math;
//   * the thread's result is one contiguous 8-byte piece (1-byte types, NB = 32) of a row, and the four
//     lanes of a quad cover a contiguous 32-byte run (128 bytes for 4-byte types).
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// CH = channels per thread (8: block of 32 columns, 4: block of 16 columns); fills 2 * CH registers
template <int CH>
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, uint32_t* r) {
  if constexpr (CH == 16) tmem_ld_16x256b_x8(taddr, r);
  else if constexpr (CH == 8) tmem_ld_16x256b_x4(taddr, r);
  else tmem_ld_16x256b_x2(taddr, r);
}

// registers -> TMEM, 32 lanes x 8 columns (thread = lane); used to pre-seed the conv1 accumulators
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// registers -> TMEM, 16 lanes x 16 columns (row-pair fragment layout, see tmem_ld_16x256b_x2)
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t* w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t w) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, const uint32_t* w) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(w[0]), "r"(w[1]) : "memory");
}
// 8-byte store executed only when flag >= 0
__device__ __forceinline__ void sts64_if(uint32_t addr, const uint32_t* w, int flag) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1, %2};\n\t}\n" ::"r"(addr), "r"(w[0]), "r"(w[1]),
               "r"(flag)
               : "memory");
}
// global stores executed only when `on`: predicated in the asm so that the compiler cannot sink the arithmetic
// that feeds them into a divergent branch
__device__ __forceinline__ void stg64_if(void* ptr, const uint32_t* w, bool on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.b32 [%0], {%1, %2};\n\t}\n" ::"l"(ptr), "r"(w[0]), "r"(w[1]),
               "r"((int)on)
               : "memory");
}
__device__ __forceinline__ void stg256_if(void* ptr, const uint32_t* w, bool on) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %9, 0;\n\t@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t@p st.global.v4.b32 [%0+16], {%5, %6, %7, %8};\n\t}\n" ::"l"(ptr),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"((int)on)
      : "memory");
}
// the same 32 bytes as ONE 256-bit store (sm_100: st.global.v8.b32, SASS STG.256; the address must be 32-byte
// aligned): a quad then writes a whole 128-byte line per instruction, and the 4-byte destinations -- whose stores
// are what the LSU is busiest with -- need half the store instructions
__device__ __forceinline__ void stg256_wide_if(void* ptr, const uint32_t* w, bool on) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %9, 0;\n\t@p st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n\t}\n" ::"l"(ptr),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"((int)on)
      : "memory");
}
// CH packed bytes (CH / 4 words) -> shared memory
template <int CH>
__device__ __forceinline__ void sts_bytes(uint32_t addr, const uint32_t* packed) {
  if constexpr (CH == 16) sts128(addr, packed);
  else if constexpr (CH == 8) sts64(addr, packed);
  else sts32(addr, packed[0]);
}

// conv0: CH accumulators (channel order) of one row -> CH bytes (CH / 4 packed words)
template <bool kDown, bool kNanSafe, int CH>
__device__ __forceinline__ void finish_conv0(const uint32_t* v, const float4* b4, const float4* s4, uint32_t* packed) {
#pragma unroll
  for (int g = 0; g < CH / 4; ++g) packed[g] = requant_u8x4<kDown, kNanSafe>(v + 4 * g, b4[g], s4[g]);
}

// conv1: CH accumulators (channel order) of one row -> CH / 4 packed words (1-byte destinations) or CH
// 32-bit words (jit_conv_kernel.cc:89-130).  `fast` selects the offset-magic conversion (c4 then holds
// C[q]; K comes per channel from k4 or, kUniK, as the one uniform `k_uni`).
template <int kDst, bool kDown, bool kNanSafe, int CH, bool kUniK>
__device__ __forceinline__ void finish_conv1(const uint32_t* v, const float4* c4, const float4* s4, const int4* k4, int k_uni,
                                             bool fast, bool relu, uint32_t* w, const float* res = nullptr) {
#pragma unroll
  for (int g = 0; g < CH / 4; ++g) {
    float t[4];
    if (fast) scale4_fast(v + 4 * g, kUniK ? make_int4(k_uni, k_uni, k_uni, k_uni) : k4[g], c4[g], s4[g], t);
    else scale4(v + 4 * g, c4[g], s4[g], t);
    if (res) {  // eltwise sum: one more separately rounded add (Params::res)
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = __fadd_rn(t[i], res[4 * g + i]);
    }
    if (kDst == DF_U8) {
      w[g] = pack_u8x4<kDown, kNanSafe>(t);
    } else if (kDst == DF_S8) {
      w[g] = pack_s8x4<kDown>(t, relu);
    } else {
      if (relu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = relu_x86(t[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) w[4 * g + i] = (kDst == DF_F32) ? __float_as_uint(t[i]) : (uint32_t)cvt_x86<kDown>(t[i]);
    }
  }
}

// CH residual values of one destination row (type / layout of the destination) as f32: u8 / s8 extend exactly,
// s32 converts like vcvtdq2ps, f32 as is
template <int kDst, int CH>
__device__ __forceinline__ void load_residual(const void* res, size_t elem, float* r) {
  if constexpr (kDst == DF_F32 || kDst == DF_S32) {
    const uint4* q = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(res) + elem * 4);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
      const uint4 x = q[i];
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) r[4 * i + e] = (kDst == DF_F32) ? __uint_as_float(xs[e]) : __int2float_rn((int)xs[e]);
    }
  } else {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(res) + elem);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
      const uint32_t x = q[i];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t b = (x >> (8 * e)) & 0xffu;
        r[4 * i + e] = (kDst == DF_U8) ? (float)b : (float)(int)(int8_t)b;
      }
    }
  }
}

// ------------------------------------------------------------------------------ epilogue role
// Work units: E0(t) = conv0 epilogue of tile t (TMEM acc0 -> u8 tile in smem) and C_j(t) = conv1 chunk j
// of tile t (TMEM acc1 -> destination).  The 16 epilogue warps form kEpiGroups groups; the warps of a group
// work on the same unit (warp w reads TMEM lane quarter w % 4 and every kBlockStride-th column block) and
// unit u of the stream belongs to group u % kEpiGroups, each group taking its units in stream order
// (kEpiGroups = 1 by default: see DF_EPI_GROUPS).  The stream order is fixed,
//     E0(0) | C_0(t) .. C_{n-2}(t)  E0(t+1)  C_{n-1}(t) | ...
// which keeps the single intermediate buffer and the two conv1 accumulators busy without ever making
// the tensor pipe wait for the epilogue it feeds: E0(t+1) starts when GEMM2(t) has read the intermediate
// tile for the last time (its last chunk was issued when C_{n-3}(t) released an accumulator), and while the
// epilogue works on C_{n-1}(t), GEMM2(t+1) already fills the other accumulator.  Tile of local iteration
// `it` = tile0 + it * tile_stride.  kPair: the barriers the MMA thread waits on live in the leader CTA.
//
// The epilogue is bound by instruction issue (16 warps on 4 schedulers; ~2.5 instructions per element
// are the floor: IADD/I2F, half a packed FADD2, half a packed FMUL2, half an F2IP), so everything that
// is not per-element work is hoisted: shared-memory addresses are 32-bit and advanced by constants, the
// static geometries always stage 1-byte output and always use ONE offset-magic constant K.
template <class G, int kDst, bool kDown0, bool kDown1, bool kNanSafe, bool kPair, class Bar>
__device__ __forceinline__ void epilogue_role(const Params& p, uint8_t* smem, Bar* bar, uint32_t tmem, int warp, int lane,
                                              int n_local, int tile0, int tile_stride) {
  const Geo<G> g{p};
  const uint32_t sbase = smem_u32(smem);
  const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
  const int group = (warp - kEpiWarp0) / kUnitWarps;            // which units this warp works on
  constexpr int kBlockStride = kUnitWarps / 4;                   // warps per lane quarter inside a group
  const int cbi = ((warp - kEpiWarp0) % kUnitWarps) >> 2;       // this warp takes column blocks cbi, cbi + kBlockStride, ...
  const int m4 = lane & 3, r8 = lane >> 2;      // position inside the row-pair fragment
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  const float* sb0 = reinterpret_cast<const float*>(smem + p.off_bias0);
  const float* ss0 = reinterpret_cast<const float*>(smem + p.off_scale0);
  const float* sb1 = reinterpret_cast<const float*>(smem + p.off_bias1);
  const float* ss1 = reinterpret_cast<const float*>(smem + p.off_scale1);
  const int* sk1 = reinterpret_cast<const int*>(smem + p.off_k1);
  constexpr int ts = (kDst == DF_F32 || kDst == DF_S32) ? 4 : 1;
  const uint32_t swz_mask1 = (uint32_t)(g.swb1() / 16 - 1);
  const bool relu1 = p.relu1 != 0;
  const bool fast1 = G::is_static ? true : (p.fast1 != 0);
  const int k_uni = p.k1_uniform;  // static geometries: never 0 (df_conv_create)
  const bool c0_only = !G::is_static && p.conv0_only != 0;
  const int q_first = p.q_first;
  // staged output (1-byte destinations, see store_staged_chunk): always for the static geometries
  constexpr bool kCanStage = (kDst == DF_U8 || kDst == DF_S8);
  const bool staged = kCanStage && !G::is_static && p.stage_out != 0;  // static geometries store straight from registers
  // where this thread's four rows (ri = 2 * h16 + hl -> tile row quarter * 32 + 8 * ri + r8) of the
  // current tile go: staging row (staged) or NHW pixel index (direct); -1 for padding positions
  int row_it = -1, pos_it = 0, rinfo[4];  // rinfo is valid for tile row_it; the PosStates are at tile pos_it
  const int m_row0 = quarter * 32 + r8;
  // position bookkeeping is spread over the warp: lane l tracks tile row quarter * 32 + l and the four
  // rows a thread needs come by shuffle from lanes r8, r8 + 8, r8 + 16, r8 + 24
  PosState pos_tile = pos_of(p, q_first + tile0 * kTileM);                         // first position of the current tile
  PosState pos_lane = pos_of(p, q_first + tile0 * kTileM + quarter * 32 + lane);   // this lane's row of it

  // arrive targets (buffer i of a pair of barriers sits 8 bytes after buffer 0)
  uint32_t a_acc0_empty = smem_u32(&bar->acc0_empty[0]), a_acc1_empty = smem_u32(&bar->acc1_empty[0]);
  uint32_t a_mid_full = smem_u32(&bar->mid_full[0]);
  if constexpr (kPair) {
    a_acc0_empty = mapa_u32(a_acc0_empty, 0);
    a_acc1_empty = mapa_u32(a_acc1_empty, 0);
    a_mid_full = mapa_u32(a_mid_full, 0);
  }
  auto arrive = [&](uint32_t a) {
    if constexpr (kPair) mbar_arrive_cluster(a);
    else mbar_arrive(a);
  };
  EpiTracer tr(p, 3, threadIdx.x == kEpiWarp0 * 32);
  if (!staged) griddep_wait();  // direct stores: earlier kernels may still be using the destination

  // Pre-seeded conv1 accumulators (static geometries: one offset-magic constant K for all channels).  The
  // epilogue warps keep the conv1 accumulators initialised to K and GEMM2 ACCUMULATES on top of it, so the
  // accumulator the epilogue reads already is the bit pattern of 2^23 + (acc - lo) -- the per-element
  // integer add of scale4_fast disappears (a third of the conv1 arithmetic instructions).  Every warp
  // (re)seeds the columns it reads, right after its tcgen05.ld has landed and before it hands the
  // accumulator back: tcgen05.st -> tcgen05.wait::st -> fence -> arrive on acc1_empty.  TMEM writes run at
  // four times the read rate (B300_MICROARCH.md), the eight registers holding K stay live for the whole role.
  constexpr bool kSeed = epilogue_seeds<G>();
  [[maybe_unused]] uint32_t kseed[8];
  auto seed_block = [&](uint32_t taddr32) {  // 32 lanes (this warp's quarter) x 32 columns at taddr32
#pragma unroll
    for (int i = 0; i < 4; ++i) tmem_st_32x32b_x8(taddr32 + 8 * i, kseed);
  };
  if constexpr (kSeed) {
#pragma unroll
    for (int i = 0; i < 8; ++i)  // volatile loads: ptxas must keep eight live registers instead of re-creating them per store
      asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(kseed[i]) : "r"(smem_u32(sk1) + 4 * i));
    for (int cb = 0; cb < 2; ++cb)
      for (int b = cbi; b < kAcc1Stride / 32; b += kBlockStride) seed_block(lane_addr + kAcc1Col + cb * kAcc1Stride + b * 32);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      arrive(a_acc1_empty);
      arrive(a_acc1_empty + 8);
    }
  }
  const int k_add = seeded_acc1<G>() ? 0 : k_uni;  // what the epilogue still has to add to the accumulator

  // ---- conv0 epilogue of local tile `it`
  auto unit_e0 = [&](int it) __attribute__((always_inline)) {
    const int ab = it % g.n_acc0();
    const int mb = it % p.NM;  // intermediate tile buffer
    mbar_wait_warp(smem_u32(&bar->mid_empty[mb]), ((it / p.NM) & 1) ^ 1);
    mbar_wait_warp(smem_u32(&bar->acc0_full[ab]), (it / g.n_acc0()) & 1);
    tc_fence_after_sync();
    tr.ev(30);
    const uint32_t mid = sbase + p.off_mid + mb * p.mid_bytes;
    const uint32_t t_base = lane_addr + ab * g.OC();
    const int nb32 = g.OC() / 32, nblk = nb32 + (g.OC() - nb32 * 32) / 16;
    auto block = [&](auto ch_c, int col0, bool last) __attribute__((always_inline)) {
      constexpr int CH = decltype(ch_c)::value;
      const int ch0 = col0 + CH * m4;  // this thread's CH consecutive conv0 channels
      uint32_t acc[2][2 * CH];
      tmem_ld_frag<CH>(t_base + col0, acc[0]);
      tmem_ld_frag<CH>(t_base + (16u << 16) + col0, acc[1]);
      float4 b4[CH / 4], s4[CH / 4];
#pragma unroll
      for (int i = 0; i < CH / 4; ++i) {
        b4[i] = *reinterpret_cast<const float4*>(sb0 + ch0 + 4 * i);
        s4[i] = *reinterpret_cast<const float4*>(ss0 + ch0 + 4 * i);
      }
      const int kb = ch0 / g.swb1();
      // byte offset of (row m_row0, channel ch0) inside the K-block, before the swizzle
      const uint32_t off0 = (uint32_t)m_row0 * g.swb1() + (uint32_t)(ch0 - kb * g.swb1());
      const uint32_t mid_kb = mid + kb * g.mid_kb_stride();
      tmem_ld_wait();
      if (last) {  // accumulator is in registers (tcgen05.wait::ld is warp-wide): the tensor pipe may overwrite it
        tc_fence_before_sync();
        if (lane == 0) arrive(a_acc0_empty + 8 * ab);
      }
      if (!dbg_flag(p, 128))
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        uint32_t v[CH], packed[CH / 4];
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = acc[ri >> 1][4 * (i / 2) + 2 * (ri & 1) + (i & 1)];
        uint32_t off = off0 + (uint32_t)(ri * 8) * g.swb1();
        off ^= ((off >> 7) & swz_mask1) << 4;  // Swizzle<B,4,3> on the (1024 B aligned) block offset
        finish_conv0<kDown0, kNanSafe, CH>(v, b4, s4, packed);
        sts_bytes<CH>(mid_kb + off, packed);
      }
    };
    bool released = false;
    for (int b = cbi; b < nblk; b += kBlockStride) {
      const bool last = b + kBlockStride >= nblk;
      released |= last;
      if (b < nb32) block(std::integral_constant<int, 8>{}, b * 32, last);
      else block(std::integral_constant<int, 4>{}, nb32 * 32, last);
    }
    if (!released) {  // warps without a block in this unit
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(a_acc0_empty + 8 * ab);
    }
    fence_proxy_async_smem();  // intermediate tile -> visible to the tensor pipe (async proxy)
    __syncwarp();
    if (lane == 0) arrive(a_mid_full + 8 * mb);
    tr.ev(31);
  };

  // ---- conv1 chunk j of local tile `it` (c = global chunk counter of this CTA)
  auto unit_c = [&](int it, int j, uint32_t c) __attribute__((always_inline)) {
    if (row_it != it) {  // tiles come in increasing order (a group may have no conv1 unit in some tile)
      for (; pos_it < it; ++pos_it) {
        pos_step(p, pos_tile, p.ts_dw, p.ts_dn, p.ts_dh);
        pos_step(p, pos_lane, p.ts_dw, p.ts_dn, p.ts_dh);
      }
      row_it = it;
      int rr = pos_pixel(p, pos_lane);
      if (rr >= 0 && staged) {
        rr -= pos_valid_before(p, pos_tile);
        // byte offset of the staging row's 16-byte unit 0 after the swizzle, row * 128 + ((row & 7) << 4);
        // unit u of the row then sits at this value XOR (u << 4)
        rr = rr * 128 + ((rr & 7) << 4);
      }
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) rinfo[ri] = __shfl_sync(0xffffffffu, rr, r8 + 8 * ri);
    }
    const int cb = c & 1;
    tr.ev(36);
    // conv0-only operator: the chunk is 128 columns of the 3x3 accumulator of this tile
    const int ab0 = it % g.n_acc0();
    if (!c0_only) { if (!dbg_flag(p, 32)) mbar_wait_warp(smem_u32(&bar->acc1_full[cb]), (c >> 1) & 1); }
    else if (j == 0) mbar_wait_warp(smem_u32(&bar->acc0_full[ab0]), (it / g.n_acc0()) & 1);
    tc_fence_after_sync();
    tr.ev(32);
    bool stage_checked = !staged;
    // which barrier tells the tensor pipe that the accumulator may be overwritten, and is it this chunk's turn
    const uint32_t a_release = c0_only ? a_acc0_empty + 8 * ab0 : a_acc1_empty + 8 * cb;
    const bool releases = !c0_only || j == g.n_chunks() - 1;
    int ncols = g.OC1() - j * g.nc1();  // real columns in this chunk
    if (ncols > g.nc1()) ncols = g.nc1();
    const uint32_t t_base = c0_only ? lane_addr + ab0 * g.OC() + j * g.nc1() : lane_addr + kAcc1Col + cb * kAcc1Stride;
    const uint32_t stage_buf = sbase + p.off_stage + cb * kStageBytes;
    const int nb32 = ncols / 32, nblk = nb32 + (ncols - nb32 * 32) / 16;
    auto block = [&](auto ch_c, auto unik_c, int col0, bool last) __attribute__((always_inline)) {
      constexpr int CH = decltype(ch_c)::value;
      constexpr bool kUniK = decltype(unik_c)::value;
      const int ccol = col0 + CH * m4;      // first of this thread's CH channels inside the chunk
      const int ch0 = j * g.nc1() + ccol;   // ... and as conv1 output channel
      uint32_t acc[2][2 * CH];
      if (!dbg_flag(p, 8)) {
        tmem_ld_frag<CH>(t_base + col0, acc[0]);
        tmem_ld_frag<CH>(t_base + (16u << 16) + col0, acc[1]);
      } else {
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) acc[0][i] = acc[1][i] = (uint32_t)(lane + i);
      }
      float4 c4[CH / 4], s4[CH / 4];
      int4 k4[kUniK ? 1 : CH / 4];
#pragma unroll
      for (int i = 0; i < CH / 4; ++i) {
        c4[i] = *reinterpret_cast<const float4*>(sb1 + ch0 + 4 * i);
        s4[i] = *reinterpret_cast<const float4*>(ss1 + ch0 + 4 * i);
        if constexpr (!kUniK) k4[i] = *reinterpret_cast<const int4*>(sk1 + ch0 + 4 * i);
      }
      // where the four rows go, computed while the TMEM loads are in flight and pinned in registers (left
      // to itself the compiler re-derives every address from the kernel parameters inside the row loop,
      // and the epilogue is bound by instruction issue)
      [[maybe_unused]] uint32_t saddr[4];
      if constexpr (ts == 1) {
        const uint32_t unit_x = ((uint32_t)ccol >> 4) << 4, stage_col = stage_buf + ((uint32_t)ccol & 15);
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
          saddr[ri] = stage_col + ((uint32_t)rinfo[ri] ^ unit_x);
          asm volatile("" : "+r"(saddr[ri]));
        }
      }
      if (!stage_checked) {  // chunk c - 2 must have left this staging buffer (waited for under the TMEM loads)
        mbar_wait_warp(smem_u32(&bar->stage_empty[cb]), ((c >> 1) & 1) ^ 1);
        stage_checked = true;
      }
      tr.ev(37);
      tmem_ld_wait();
      tr.ev(38);
      if constexpr (kSeed) seed_block(t_base + col0);  // the columns just read are K again for the chunk after next
      if (last && releases) {  // accumulator is in registers (tcgen05.wait::ld is warp-wide): the tensor pipe may overwrite it
        if constexpr (kSeed) tmem_st_wait();
        tc_fence_before_sync();
        if (lane == 0) arrive(a_release);
      }
      tr.ev(39);
      // Padding rows are computed like any other and only their store is predicated off: a branch around
      // the row costs three control instructions and a branch-resolve stall per row, the wasted arithmetic
      // (6..8 % of the rows for the BASELINE shapes) is cheaper.
      if (!dbg_flag(p, 16))
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        const int rr = rinfo[ri];
        uint32_t v[CH], w[ts == 1 ? CH / 4 : CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = acc[ri >> 1][4 * (i / 2) + 2 * (ri & 1) + (i & 1)];
        if (p.res != nullptr) {  // eltwise sum (never staged: rr is the pixel index)
          float rs[CH];
          if (rr >= 0) load_residual<kDst, CH>(p.res, (size_t)rr * p.dst_pitch + p.dst_ch0 + ch0, rs);
          else {
#pragma unroll
            for (int i = 0; i < CH; ++i) rs[i] = 0.f;
          }
          finish_conv1<kDst, kDown1, kNanSafe, CH, kUniK>(v, c4, s4, k4, k_add, fast1, relu1, w, rs);
        } else {
          finish_conv1<kDst, kDown1, kNanSafe, CH, kUniK>(v, c4, s4, k4, k_add, fast1, relu1, w);
        }
        if constexpr (ts == 1) {
          if (staged) {  // 16-byte unit XOR row-inside-the-1024-B-atom, as SWIZZLE_128B wants
            if (rr >= 0 && !dbg_flag(p, 4)) sts_bytes<CH>(saddr[ri], w);
          } else if (rr >= 0 && !dbg_flag(p, 4)) {
            uint8_t* out = static_cast<uint8_t*>(p.dst) + (size_t)rr * p.dst_pitch + p.dst_ch0 + ch0;
            if constexpr (CH == 8) *reinterpret_cast<uint2*>(out) = make_uint2(w[0], w[1]);
            else *reinterpret_cast<uint32_t*>(out) = w[0];
          }
        } else if (rr >= 0) {
          uint4* out = reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.dst) + ((size_t)rr * p.dst_pitch + p.dst_ch0 + ch0) * 4);
          if (CH == 8 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) {
            stg256_wide_if(out, w, true);  // one 256-bit store (see stg256_wide_if)
          } else {
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) out[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          }
        }
      }
    };
    bool released = false;
    for (int b = cbi; b < nblk; b += kBlockStride) {
      const bool last = b + kBlockStride >= nblk;
      released |= last;
      if (G::is_static || k_uni != 0) {
        if (b < nb32) block(std::integral_constant<int, 8>{}, std::true_type{}, b * 32, last);
        else block(std::integral_constant<int, 4>{}, std::true_type{}, nb32 * 32, last);
      } else {
        if (b < nb32) block(std::integral_constant<int, 8>{}, std::false_type{}, b * 32, last);
        else block(std::integral_constant<int, 4>{}, std::false_type{}, nb32 * 32, last);
      }
    }
    if (!released && releases) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(a_release);
    }
    tr.ev(35);
    // warps without a block in this chunk (ragged last chunk) must not report the staging buffer full
    // before chunk c - 2 has left it either: their arrival completes the phase the store thread waits for
    if (!stage_checked) mbar_wait_warp(smem_u32(&bar->stage_empty[cb]), ((c >> 1) & 1) ^ 1);
    if (staged) {
      fence_proxy_async_smem();  // this thread's staging writes -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar->stage_full[cb]));
    }
    tr.ev(33);
  };

  // ---- the unit stream (see above): every warp walks it, a group executes its own units
  const int nch = g.n_chunks();
  uint32_t c = 0, u = 0;
  auto mine = [&]() { return (u++ % kEpiGroups) == (uint32_t)group; };  // unit u belongs to group u % kEpiGroups
  if (c0_only) {  // conv0-only operator: no intermediate tile, every unit finishes accumulator columns;
                  // all chunks of a tile stay with one group (they share the tile's accumulator hand-off)
    for (int it = 0; it < n_local; ++it)
      for (int j = 0; j < nch; ++j, ++c)
        if ((it % kEpiGroups) == group) unit_c(it, j, c);
    return;
  }
  if (n_local > 0 && mine()) unit_e0(0);
  for (int it = 0; it < n_local; ++it) {
    for (int j = 0; j < nch; ++j, ++c) {
      if (j == nch - 1 && it + 1 < n_local && mine()) unit_e0(it + 1);
      if (mine()) unit_c(it, j, c);
    }
  }
}

template <int N, class F, int I = 0>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<N, F, I + 1>(static_cast<F&&>(f));
  }
}

// ------------------------------------------------------------------ epilogue role, static geometries
// The BASELINE shapes (StaticGeom: OC a multiple of 32, conv1 chunks of 128 channels, finite constants,
// round-to-nearest, one offset-magic constant K) run this epilogue instead of epilogue_role.  Same unit
// stream, same hand-offs, but written for the one thing that bounds it: instruction issue.  The 16 epilogue
// warps share four schedulers with the single-thread roles, and profiles/r02_skeleton_ncu.txt shows that of
// the ~17.6 k warp instructions epilogue_role executes per 128-position tile only ~5 k are arithmetic: the rest
// is per-unit bookkeeping (addresses re-derived from kernel parameters, staging hand-offs, polling loops), and
// with every unit knocked down to its skeleton the kernel is only half as fast again as the real one.  Here
//   * output goes straight from registers to global memory: a thread owns 8 consecutive channels of a row,
//     a quad 32 contiguous bytes (128 for 4-byte types) -- no staging buffers, store thread, proxy fences
//     or staging barriers (measured faster than both TMA-store variants, profiles/r02_variants.log);
//   * everything a conv1 unit needs besides its accumulator is computed once per tile (row -> pixel index
//     from one ballot over the warp's 32 rows: valid positions are a contiguous pixel range) or once per
//     kernel (shared-memory addresses), and the chunk index only moves constant offsets;
//   * optionally (DF_SEED) the conv1 accumulators are kept pre-seeded with K so that GEMM2 accumulates on
//     top of it and the per-element integer add of the offset-magic conversion disappears.
template <class G, int kDst, bool kPair, class Bar>
__device__ __forceinline__ void epilogue_static(const Params& p, const DstMaps&, uint8_t* smem, Bar* bar, uint32_t tmem,
                                                int warp, int lane, int n_local, int tile0, int tile_stride) {
  static_assert(G::is_static && G::nc1 == 128 && G::OC % 32 == 0, "epilogue_static: unsupported geometry");
  const uint32_t sbase = smem_u32(smem);
  const int e = warp - kEpiWarp0;               // 0 .. 15
  const int quarter = e & 3;                    // TMEM lane quarter (= warp & 3)
  const int cbi = e >> 2;                       // 32-column block of a 128-column chunk / of the conv0 accumulator
  const int m4 = lane & 3, r8 = lane >> 2;
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  constexpr int ts = (kDst == DF_F32 || kDst == DF_S32) ? 4 : 1;
  const bool relu1 = p.relu1 != 0;
  const int q_first = p.q_first;
  // this thread's 8 channels inside a chunk / a conv0 block row, and where their constants live
  const int chl = cbi * 32 + 8 * m4;
  const uint32_t sb1 = sbase + p.off_bias1 + 4 * chl, ss1 = sbase + p.off_scale1 + 4 * chl;
  const uint32_t bar_acc1_full = smem_u32(&bar->acc1_full[0]), bar_acc0_full = smem_u32(&bar->acc0_full[0]);
  const uint32_t bar_mid_empty = smem_u32(&bar->mid_empty[0]);
  uint32_t a_acc0_empty = smem_u32(&bar->acc0_empty[0]), a_acc1_empty = smem_u32(&bar->acc1_empty[0]);
  uint32_t a_mid_full = smem_u32(&bar->mid_full[0]);
  if constexpr (kPair) {
    a_acc0_empty = mapa_u32(a_acc0_empty, 0);
    a_acc1_empty = mapa_u32(a_acc1_empty, 0);
    a_mid_full = mapa_u32(a_mid_full, 0);
  }
  auto arrive = [&](uint32_t a) {
    if constexpr (kPair) mbar_arrive_cluster(a);
    else mbar_arrive(a);
  };
  EpiTracer tr(p, 3, threadIdx.x == kEpiWarp0 * 32);
  griddep_wait();  // earlier kernels in the stream may still be using the destination

  // ---- pre-seeded conv1 accumulators: this warp's 32 lanes x 32 columns of both buffers
  constexpr bool kSeed = epilogue_seeds<G>();
  const int k_add = seeded_acc1<G>() ? 0 : p.k1_uniform;
  const bool wide_st = DF_WIDE_ST && (reinterpret_cast<uintptr_t>(p.dst) & 31) == 0;  // 4-byte destinations: 256-bit stores
  [[maybe_unused]] uint32_t kseed[8];
  const uint32_t t_acc1 = lane_addr + kAcc1Col + cbi * 32;  // buffer 0; buffer 1: + kAcc1Stride
  auto seed_block = [&](uint32_t taddr32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tmem_st_32x32b_x8(taddr32 + 8 * i, kseed);
  };
  if constexpr (kSeed) {
#pragma unroll
    for (int i = 0; i < 8; ++i)  // volatile loads: ptxas must keep eight live registers instead of re-creating them per store
      asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(kseed[i]) : "r"(sbase + p.off_k1 + 4 * i));
    seed_block(t_acc1);
    seed_block(t_acc1 + kAcc1Stride);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      arrive(a_acc1_empty);
      arrive(a_acc1_empty + 8);
    }
  }

  // ---- position bookkeeping: lane l tracks tile row quarter * 32 + l (see PosState); the valid rows of the
  //      warp's 32 are the contiguous pixel range [f0, f0 + popc(mask)), row r is pixel f0 + popc(mask below r)
  PosState pos_lane = pos_of(p, q_first + tile0 * kTileM + quarter * 32 + lane);
  // pixel index of this lane's row in the NEXT tile tile_rows() will be asked for, computed one tile ahead: the
  // carries and constant-bank loads behind it are a ~500-cycle dependent chain, which sat between the last unit of a
  // tile and the first of the next (profiles/r02_variants_tile_rows.log); issued a tile early, nobody waits for it
  int pix_next = pos_pixel(p, pos_lane);
  uint8_t* rptr[4];  // where the thread's rows r8 + 8 ri go (its 8 channels of chunk 0)
  uint32_t rvalid = 0;  // bit ri: row is a real pixel.  (Sending padding rows to a scratch area instead of predicating
                        // the stores was 3.5x slower: every SM hammering the same few L2 lines.)
  auto tile_rows = [&](int) {  // called once per local tile, in order
    const int pix = pix_next;
    pos_step(p, pos_lane, p.ts_dw, p.ts_dn, p.ts_dh);
    pix_next = pos_pixel(p, pos_lane);
    const uint32_t mask = __ballot_sync(0xffffffffu, pix >= 0);
    const int f0 = __shfl_sync(0xffffffffu, pix, mask ? __ffs(mask) - 1 : 0);
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
      const int r = r8 + 8 * ri;
      const int rho = __popc(mask & ((1u << r) - 1u));
      rptr[ri] = static_cast<uint8_t*>(p.dst) + ((size_t)(f0 + rho) * G::OC1 + chl) * ts;
      rvalid = (rvalid & ~(1u << ri)) | (((mask >> r) & 1u) << ri);
    }
  };

  // ---- conv0 epilogue of local tile `it`: acc0 -> u8 -> intermediate tile in smem (K-major, swizzled);
  //      warp e takes 32 rows (its lane quarter) x the 32-column blocks cbi, cbi + 4, ...
  constexpr uint32_t swz_mask1 = (uint32_t)(G::swb1 / 16 - 1);
  constexpr int nb0 = G::OC / 32;
  constexpr int kbw = G::swb1;
  auto unit_e0 = [&](int it) {
    const int ab = it % G::n_acc0;
    const int mb = it % p.NM;  // intermediate tile buffer
    mbar_wait_warp(bar_mid_empty + 8 * mb, ((it / p.NM) & 1) ^ 1);
    mbar_wait_warp(bar_acc0_full + 8 * ab, (it / G::n_acc0) & 1);
    tc_fence_after_sync();
    tr.ev(30);
    const uint32_t mid = sbase + p.off_mid + mb * p.mid_bytes;
    const uint32_t t_base = lane_addr + ab * G::OC;
    bool released = false;
#pragma unroll
    for (int b = cbi; b < nb0; b += 4) {
      const int ch0 = b * 32 + 8 * m4;
      uint32_t acc[2][16];
      tmem_ld_frag<8>(t_base + b * 32, acc[0]);
      tmem_ld_frag<8>(t_base + (16u << 16) + b * 32, acc[1]);
      float4 b4[2], s4[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        b4[i] = lds128f(sbase + p.off_bias0 + 4 * (ch0 + 4 * i));
        s4[i] = lds128f(sbase + p.off_scale0 + 4 * (ch0 + 4 * i));
      }
      const int kb = ch0 / kbw;
      const uint32_t off0 = (uint32_t)(quarter * 32 + r8) * kbw + (uint32_t)(ch0 - kb * kbw);
      const uint32_t mid_kb = mid + kb * (uint32_t)(kTileM * kbw);
      tmem_ld_wait();
      if (b + 4 >= nb0) {  // last block of this warp: the accumulator is in registers
        tc_fence_before_sync();
        if (lane == 0) arrive(a_acc0_empty + 8 * ab);
        released = true;
      }
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        uint32_t v[8], packed[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[ri >> 1][4 * (i / 2) + 2 * (ri & 1) + (i & 1)];
        uint32_t off = off0 + (uint32_t)(ri * 8) * kbw;
        off ^= ((off >> 7) & swz_mask1) << 4;
        finish_conv0<false, false, 8>(v, b4, s4, packed);
        sts_bytes<8>(mid_kb + off, packed);
      }
    }
    if (!released) {  // warps without a conv0 block (OC < 128)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(a_acc0_empty + 8 * ab);
    }
    tr.ev(34);
    fence_proxy_async_smem();  // intermediate tile -> visible to the tensor pipe (async proxy)
    __syncwarp();
    if (lane == 0) arrive(a_mid_full + 8 * mb);
    tr.ev(31);
  };

  // ---- conv1 chunk j of the current tile (c = chunk counter of this CTA): this warp's 32 rows x 32 channels
  // A barrier test costs ~150 cycles even when the phase is complete (B300_MICROARCH.md: test_wait 149), and in
  // this lock-step stream nothing else runs meanwhile.  So every unit tests the NEXT unit's accumulator barrier
  // before its own arithmetic (`pre`, consumed a unit later): in steady state the test's latency then sits
  // under ~50 arithmetic instructions and the unit starts without waiting.
  uint32_t pre = 0;  // result of the early test of chunk c's barrier (all lanes agree)
  auto unit_c = [&](auto j_c, uint32_t c) {
    constexpr int j = decltype(j_c)::value;  // compile-time chunk index: constant offsets everywhere
    const uint32_t cb = c & 1;
    tr.ev(36);
    if (!__all_sync(0xffffffffu, pre != 0)) mbar_wait_warp(bar_acc1_full + 8 * cb, (c >> 1) & 1);
    tc_fence_after_sync();
    tr.ev(32);
    uint32_t acc[2][16];
    const uint32_t t_blk = t_acc1 + cb * kAcc1Stride;
    tmem_ld_frag<8>(t_blk, acc[0]);
    tmem_ld_frag<8>(t_blk + (16u << 16), acc[1]);
    float4 c4[2], s4[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      c4[i] = lds128f(sb1 + 4 * (j * 128 + 4 * i));
      s4[i] = lds128f(ss1 + 4 * (j * 128 + 4 * i));
    }
    tmem_ld_wait();
    tr.ev(38);
    if constexpr (kSeed) {
      seed_block(t_blk);  // the columns just read are K again for the chunk after next
      tmem_st_wait();
    }
    tc_fence_before_sync();
    if (lane == 0) arrive(a_acc1_empty + 8 * cb);  // accumulator is in registers: the tensor pipe may overwrite it
    tr.ev(39);
    pre = mbar_test_wait(bar_acc1_full + 8 * (cb ^ 1), ((c + 1) >> 1) & 1) ? 1u : 0u;  // chunk c + 1, used by the next unit
    // Padding rows are computed like any other and only their store is predicated off (a branch around a
    // row costs more than the wasted arithmetic: 6..8 % of the rows for the BASELINE shapes).
    if (!dbg_flag(p, 16))
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
      uint32_t v[8], w[ts == 1 ? 2 : 8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = acc[ri >> 1][4 * (i / 2) + 2 * (ri & 1) + (i & 1)];
      finish_conv1<kDst, false, false, 8, true>(v, c4, s4, nullptr, k_add, true, relu1, w);
      uint8_t* out = rptr[ri] + j * (128 * ts);
      if constexpr (ts == 1) stg64_if(out, w, ((rvalid >> ri) & 1u) && !dbg_flag(p, 4));
      else if (wide_st) stg256_wide_if(out, w, ((rvalid >> ri) & 1u) && !dbg_flag(p, 4));
      else stg256_if(out, w, ((rvalid >> ri) & 1u) && !dbg_flag(p, 4));
    }
    tr.ev(33);
  };

  // ---- the unit stream:  E0(0) | C_0(t) .. E0(t+1) C_{e0_pos}(t) .. C_{n-1}(t) | ...
#ifndef DF_E0_POS
#define DF_E0_POS (G::n_chunks / 2)  // measured best of the four positions for cfg3 (profiles/r02_variants_e0_position.log)
#endif
  // With ONE intermediate tile E0(t+1) has to wait until GEMM2 has read tile t for the last time, and the last
  // chunk is only issued once C_{n-3}(t) has released an accumulator: anything earlier than n - 1 deadlocks.
  const int e0_pos = p.NM >= 2 ? ((DF_E0_POS) < G::n_chunks ? (DF_E0_POS) : G::n_chunks - 1) : G::n_chunks - 1;

  // (Tried and dropped in the last session of round 2, each parity-green on cfg3 and with its log under profiles/:
  //  the unit stream as a software pipeline of half units, tcgen05.ld of half x + 1 in flight under the arithmetic of
  //  half x: 3..4 % slower, 9..14 % when the accumulator is also released after the arithmetic
  //  (r02_variants_half_pipeline.log); E0's two barrier tests issued by the unit in front of it: no change
  //  (r02_variants_e0_pretest.log); three conv1 accumulators + one conv0 accumulator in TMEM: 5 % slower
  //  (r02_variants_three_acc1_buffers.log).)
  // (Tried and dropped: moving the block in four 16 x 16 pieces through a ring of register sets with the loads
  // running two pieces ahead ACROSS units -- parity-green but 6 % slower on cfg3 and 19 % on cfg1,
  // profiles/r02_variants_piece_pipeline.log: the extra loads / waits cost more issue slots than the hidden latency
  // returned.  Likewise two groups of 8 warps on alternate accumulators with per-warp TMA stores,
  // profiles/r02_variants_seed_epilogue_design.log.)
  uint32_t c = 0;
  if (n_local > 0) unit_e0(0);
  // (Tried: tile_rows under the first chunk's TMEM load instead of between the tiles -- the row pointers are the address
  // registers of the previous tile's last stores, and overwriting them waits ~600 cycles for the LSU to take those
  // stores; inside the unit that wait delays the accumulator's release: 10 % slower on cfg1,
  // profiles/r02_variants_tile_rows.log.)
  for (int it = 0; it < n_local; ++it) {
    tile_rows(it);
    static_for<G::n_chunks>([&](auto j_c) {
      constexpr int j = decltype(j_c)::value;
      if (j == e0_pos && it + 1 < n_local) unit_e0(it + 1);
      unit_c(j_c, c);
      ++c;
    });
  }
}


// ------------------------------------------------------------- epilogue role, static geometries, two groups
// Same units, same hand-offs and the same arithmetic as epilogue_static, but the 16 warps form two groups of eight
// that never wait for each other (DF_STATIC_GROUPS == 2, n_chunks even):
//   * group g reads conv1 accumulator g: chunk c of the CTA's chunk stream goes to buffer c & 1, and with an even
//     number of chunks per tile that is chunk index j & 1 -- group 0 takes C_0, C_2, ..., group 1 takes C_1, C_3, ...;
//   * E0(t) belongs to group t & 1 and sits in that group's sequence in front of its first chunk j >= e0_pos of tile
//     t - 1 (E0(0) comes first);
//   * a warp owns its lane quarter's 32 rows x TWO 32-column blocks of a unit (blocks cbi and cbi + 2).  The two blocks
//     go through ONE set of 32 accumulator registers: the second block's TMEM loads are issued into the halves the
//     first block's arithmetic has already consumed, so their latency runs under that arithmetic; the accumulator is
//     re-seeded and released once, after both blocks have landed.
template <class G, int kDst, bool kPair, class Bar>
__device__ __forceinline__ void epilogue_static2(const Params& p, uint8_t* smem, Bar* bar, uint32_t tmem, int warp, int lane,
                                                 int n_local, int tile0, int tile_stride) {
  static_assert(G::is_static && G::nc1 == 128 && G::OC % 32 == 0 && G::n_chunks % 2 == 0, "epilogue_static2: unsupported geometry");
  const uint32_t sbase = smem_u32(smem);
  const int e = warp - kEpiWarp0;               // 0 .. 15
  const int grp = e >> 3;                       // group = conv1 accumulator buffer
  const int quarter = e & 3;                    // TMEM lane quarter (= warp & 3)
  const int cbi = (e >> 2) & 1;                 // first of this warp's 32-column blocks; the other one is cbi + 2
  const int m4 = lane & 3, r8 = lane >> 2;
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  constexpr int ts = (kDst == DF_F32 || kDst == DF_S32) ? 4 : 1;
  const bool relu1 = p.relu1 != 0;
  const int q_first = p.q_first;
  const int chl = cbi * 32 + 8 * m4;            // this thread's 8 channels inside block cbi of a chunk (+64: block cbi + 2)
  const uint32_t sb1 = sbase + p.off_bias1 + 4 * chl, ss1 = sbase + p.off_scale1 + 4 * chl;
  const uint32_t bar_acc1_full = smem_u32(&bar->acc1_full[grp]), bar_acc0_full = smem_u32(&bar->acc0_full[0]);
  const uint32_t bar_mid_empty = smem_u32(&bar->mid_empty[0]);
  uint32_t a_acc0_empty = smem_u32(&bar->acc0_empty[0]), a_acc1_empty = smem_u32(&bar->acc1_empty[grp]);
  uint32_t a_mid_full = smem_u32(&bar->mid_full[0]);
  if constexpr (kPair) {
    a_acc0_empty = mapa_u32(a_acc0_empty, 0);
    a_acc1_empty = mapa_u32(a_acc1_empty, 0);
    a_mid_full = mapa_u32(a_mid_full, 0);
  }
  auto arrive = [&](uint32_t a) {
    if constexpr (kPair) mbar_arrive_cluster(a);
    else mbar_arrive(a);
  };
  griddep_wait();  // earlier kernels in the stream may still be using the destination

  constexpr bool kSeed = epilogue_seeds<G>();
  const int k_add = seeded_acc1<G>() ? 0 : p.k1_uniform;
  const bool wide_st = DF_WIDE_ST && (reinterpret_cast<uintptr_t>(p.dst) & 31) == 0;
  [[maybe_unused]] uint32_t kseed[8];
  const uint32_t t_acc1 = lane_addr + kAcc1Col + grp * kAcc1Stride + cbi * 32;  // block cbi of this group's buffer; block cbi + 2: + 64
  auto seed_block = [&](uint32_t taddr32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tmem_st_32x32b_x8(taddr32 + 8 * i, kseed);
  };
  if constexpr (kSeed) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(kseed[i]) : "r"(sbase + p.off_k1 + 4 * i));
    seed_block(t_acc1);
    seed_block(t_acc1 + 64);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) arrive(a_acc1_empty);
  }

  PosState pos_lane = pos_of(p, q_first + tile0 * kTileM + quarter * 32 + lane);
  int pix_next = pos_pixel(p, pos_lane);
  uint8_t* rptr[4];
  uint32_t rvalid = 0;
  auto tile_rows = [&]() {  // once per local tile, in order
    const int pix = pix_next;
    pos_step(p, pos_lane, p.ts_dw, p.ts_dn, p.ts_dh);
    pix_next = pos_pixel(p, pos_lane);
    const uint32_t mask = __ballot_sync(0xffffffffu, pix >= 0);
    const int f0 = __shfl_sync(0xffffffffu, pix, mask ? __ffs(mask) - 1 : 0);
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
      const int r = r8 + 8 * ri;
      const int rho = __popc(mask & ((1u << r) - 1u));
      rptr[ri] = static_cast<uint8_t*>(p.dst) + ((size_t)(f0 + rho) * G::OC1 + chl) * ts;
      rvalid = (rvalid & ~(1u << ri)) | (((mask >> r) & 1u) << ri);
    }
  };

  // ---- conv0 epilogue of local tile `it` (this group's turn): blocks cbi, cbi + 2, ... of the conv0 accumulator
  constexpr uint32_t swz_mask1 = (uint32_t)(G::swb1 / 16 - 1);
  constexpr int nb0 = G::OC / 32;
  constexpr int kbw = G::swb1;
  auto unit_e0 = [&](int it) __attribute__((always_inline)) {
    const int ab = it % G::n_acc0;
    const int mb = it % p.NM;
    mbar_wait_warp(bar_mid_empty + 8 * mb, ((it / p.NM) & 1) ^ 1);
    mbar_wait_warp(bar_acc0_full + 8 * ab, (it / G::n_acc0) & 1);
    tc_fence_after_sync();
    const uint32_t mid = sbase + p.off_mid + mb * p.mid_bytes;
    const uint32_t t_base = lane_addr + ab * G::OC;
    bool released = false;
#pragma unroll
    for (int b = cbi; b < nb0; b += 2) {
      const int ch0 = b * 32 + 8 * m4;
      uint32_t acc[2][16];
      tmem_ld_frag<8>(t_base + b * 32, acc[0]);
      tmem_ld_frag<8>(t_base + (16u << 16) + b * 32, acc[1]);
      float4 b4[2], s4[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        b4[i] = lds128f(sbase + p.off_bias0 + 4 * (ch0 + 4 * i));
        s4[i] = lds128f(sbase + p.off_scale0 + 4 * (ch0 + 4 * i));
      }
      const int kb = ch0 / kbw;
      const uint32_t off0 = (uint32_t)(quarter * 32 + r8) * kbw + (uint32_t)(ch0 - kb * kbw);
      const uint32_t mid_kb = mid + kb * (uint32_t)(kTileM * kbw);
      tmem_ld_wait();
      if (b + 2 >= nb0) {  // last block of this warp: the accumulator is in registers
        tc_fence_before_sync();
        if (lane == 0) arrive(a_acc0_empty + 8 * ab);
        released = true;
      }
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        uint32_t v[8], packed[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[ri >> 1][4 * (i / 2) + 2 * (ri & 1) + (i & 1)];
        uint32_t off = off0 + (uint32_t)(ri * 8) * kbw;
        off ^= ((off >> 7) & swz_mask1) << 4;
        finish_conv0<false, false, 8>(v, b4, s4, packed);
        sts_bytes<8>(mid_kb + off, packed);
      }
    }
    if (!released) {  // warps without a conv0 block (OC < 64)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive(a_acc0_empty + 8 * ab);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) arrive(a_mid_full + 8 * mb);
  };

  // ---- conv1 chunk j (compile time) of the current tile; u = how many chunks this group has read before it
  auto rows2 = [&](const uint32_t* a16, int half, const float4* c4, const float4* s4, int col_off) __attribute__((always_inline)) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int ri = 2 * half + rr;
      uint32_t v[8], w[ts == 1 ? 2 : 8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = a16[4 * (i / 2) + 2 * rr + (i & 1)];
      finish_conv1<kDst, false, false, 8, true>(v, c4, s4, nullptr, k_add, true, relu1, w);
      uint8_t* out = rptr[ri] + col_off;
      if constexpr (ts == 1) stg64_if(out, w, (rvalid >> ri) & 1u);
      else if (wide_st) stg256_wide_if(out, w, (rvalid >> ri) & 1u);
      else stg256_if(out, w, (rvalid >> ri) & 1u);
    }
  };
  uint32_t pre = 0;
  auto unit_c = [&](auto j_c, uint32_t u) __attribute__((always_inline)) {
    constexpr int j = decltype(j_c)::value;
    if (!__all_sync(0xffffffffu, pre != 0)) mbar_wait_warp(bar_acc1_full, u & 1);
    tc_fence_after_sync();
    uint32_t acc[2][16];
    tmem_ld_frag<8>(t_acc1, acc[0]);
    tmem_ld_frag<8>(t_acc1 + (16u << 16), acc[1]);
    float4 c4[2], s4[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      c4[i] = lds128f(sb1 + 4 * (j * 128 + 4 * i));
      s4[i] = lds128f(ss1 + 4 * (j * 128 + 4 * i));
    }
    tmem_ld_wait();
    // block cbi: rows 0..15 -> then their registers take block cbi + 2's rows 0..15; likewise the upper rows
    rows2(acc[0], 0, c4, s4, j * (128 * ts));
    tmem_ld_frag<8>(t_acc1 + 64, acc[0]);
    rows2(acc[1], 1, c4, s4, j * (128 * ts));
    tmem_ld_frag<8>(t_acc1 + 64 + (16u << 16), acc[1]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      c4[i] = lds128f(sb1 + 4 * (j * 128 + 64 + 4 * i));
      s4[i] = lds128f(ss1 + 4 * (j * 128 + 64 + 4 * i));
    }
    tmem_ld_wait();
    if constexpr (kSeed) {
      seed_block(t_acc1);
      seed_block(t_acc1 + 64);
      tmem_st_wait();
    }
    tc_fence_before_sync();
    if (lane == 0) arrive(a_acc1_empty);  // both blocks are in registers: the tensor pipe may overwrite the accumulator
    pre = mbar_test_wait(bar_acc1_full, (u + 1) & 1) ? 1u : 0u;  // this group's next chunk
    rows2(acc[0], 0, c4, s4, j * (128 * ts) + 64 * ts);
    rows2(acc[1], 1, c4, s4, j * (128 * ts) + 64 * ts);
  };

#ifndef DF_E0_POS
#define DF_E0_POS (G::n_chunks / 2)
#endif
  const int e0_pos = p.NM >= 2 ? ((DF_E0_POS) < G::n_chunks ? (DF_E0_POS) : G::n_chunks - 1) : G::n_chunks - 1;
  uint32_t u = 0;  // chunks this group has read
  if (n_local > 0 && grp == 0) unit_e0(0);
  for (int it = 0; it < n_local; ++it) {
    tile_rows();
    bool e0_done = !(it + 1 < n_local && ((it + 1) & 1) == grp);  // is E0(it + 1) ours, and still to do?
    static_for<G::n_chunks>([&](auto j_c) {
      constexpr int j = decltype(j_c)::value;
      if ((j & 1) == grp) {
        if (!e0_done && j >= e0_pos) {
          unit_e0(it + 1);
          e0_done = true;
        }
        unit_c(j_c, u);
        ++u;
      }
    });
    if (!e0_done) unit_e0(it + 1);  // (no chunk of ours at or after e0_pos)
  }
}

// ---- store thread of the staged output path (one elected thread of warp 3): sends every staged conv1
// chunk to the destination (store_staged_chunk) in the order the epilogue produces them
template <class G, class Bar>
__device__ __forceinline__ void store_role(const Params& p, const DstMaps& tmD, Bar* bar, uint32_t sbase, int n_local,
                                           int tile0, int tile_stride) {
  const Geo<G> g{p};
  const int q_first = p.q_first;
  Tracer tr(p, 2);
  uint32_t c = 0;
  griddep_wait();  // earlier kernels in the stream may still be reading / writing the destination
  for (int it = 0; it < n_local; ++it) {
    const int q0 = q_first + (tile0 + it * tile_stride) * kTileM;
    const int f0 = valid_before(p, q0), V = valid_before(p, q0 + kTileM) - f0;
    for (int j = 0; j < g.n_chunks(); ++j, ++c) {
      const uint32_t cb = c & 1;
      mbar_wait(smem_u32(&bar->stage_full[cb]), (c >> 1) & 1);
      tr.ev(40);
      if (!dbg_flag(p, 2)) store_staged_chunk(tmD, sbase + p.off_stage + cb * kStageBytes, f0, V, p.dst_ch0 + j * g.nc1());
      tr.ev(41);
      bulk_wait_read_all();
      tr.ev(42);
      mbar_arrive(smem_u32(&bar->stage_empty[cb]));
    }
  }
  bulk_wait_all();  // the staging buffers must outlive the last TMA stores
}

// Per-channel f32 bias / scale (and offset-magic K) vectors -> shared memory, by the epilogue warps only and
// AFTER the CTA-wide start barrier: the global loads' latency (~1 us) then overlaps the first halo / weight
// loads and the first GEMM1 instead of delaying every role.
template <class G>
__device__ __forceinline__ void load_epilogue_constants(const Params& p, uint8_t* smem) {
  const Geo<G> g{p};
  float* sb0 = reinterpret_cast<float*>(smem + p.off_bias0);
  float* ss0 = reinterpret_cast<float*>(smem + p.off_scale0);
  float* sb1 = reinterpret_cast<float*>(smem + p.off_bias1);
  float* ss1 = reinterpret_cast<float*>(smem + p.off_scale1);
  int* sk1 = reinterpret_cast<int*>(smem + p.off_k1);
  const int t = (int)threadIdx.x - kEpiWarp0 * 32, nt = kEpiWarps * 32;
  for (int i = t; i < g.OC(); i += nt) {
    sb0[i] = p.bias0[i];
    ss0[i] = p.scale0[i];
  }
  const int oc1_pad = g.n_chunks() * g.nc1();
  for (int i = t; i < oc1_pad; i += nt) {
    sb1[i] = i < g.OC1() ? p.bias1[i] : 0.f;
    ss1[i] = i < g.OC1() ? p.scale1[i] : 0.f;
    sk1[i] = i < g.OC1() ? p.k1[i] : 0;
  }
  named_bar_sync(1, nt);  // epilogue warps only
}

// ------------------------------------------------------------------------------- the kernel
template <class G, int kDst, bool kDown0, bool kDown1, bool kNanSafe>
__global__ void __launch_bounds__(kThreads, 1)
conv_fused_kernel(const __grid_constant__ SrcMaps tmS, const __grid_constant__ CUtensorMap tmW0,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ DstMaps tmD,
                  const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align in the shared address space (keeps LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Barriers* bar = reinterpret_cast<Barriers*>(smem);
  const uint32_t sbase = smem_u32(smem);
  trace_wallclock(p, 0);
  // PDL: the next launch may take over SMs as soon as this grid's CTAs leave them and run its prologue
  // (barriers, TMEM, weights) under our tail; whatever depends on earlier kernels sits behind griddep_wait()
  griddep_launch_dependents();
  const Geo<G> g{p};

  // shfl from lane 0 tells ptxas the warp index is warp-uniform (values derived from it can then
  // live in uniform registers, e.g. the column base of the epilogue's constant-bank loads)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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
      mbar_init(smem_u32(&bar->acc0_empty[i]), unit_warps<G>());
      mbar_init(smem_u32(&bar->mid_full[i]), unit_warps<G>());
      mbar_init(smem_u32(&bar->mid_empty[i]), 1);
      mbar_init(smem_u32(&bar->acc1_full[i]), 1);
      mbar_init(smem_u32(&bar->acc1_empty[i]), unit_warps<G>());
      mbar_init(smem_u32(&bar->stage_full[i]), kUnitWarps);
      mbar_init(smem_u32(&bar->stage_empty[i]), 1);
    }
    for (int i = 0; i < kG1Ahead; ++i) mbar_init(smem_u32(&bar->g1_prog[i]), 1);
    for (int i = 0; i < kMaxAStages; ++i) mbar_init(smem_u32(&bar->a_ready[i]), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmS.m[0]);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
  }
  if (warp == 3) tmem_alloc<512>(smem_u32(&bar->tmem_base));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bar->tmem_base;

  const int q_first = p.q_first;  // linear index of image 0, row 0, column 0

  // The three single-thread roles each run their whole loop inside ONE elect.sync region and advance
  // shared-memory descriptors by ADDITION: that keeps descriptor math in the uniform datapath.  Both
  // alternatives measured slow (probe/mma_contention.cu, profiles/r01_mma_issue_probe.log): a
  // per-tap elect/__syncwarp costs ~370 cycles per iteration, and rebuilding descriptors from
  // vector registers (R2UR) ~140 cycles per tap -- more than the 96..256 cycles of MMA work in a tap.
  // fused concat -> conv (run-time geometry only): K-blocks come from several tensor maps, and with the concat's
  // ReLU the whole producer warp clamps every halo stage before the tensor pipe may read it
  const bool multi_src = !G::is_static && p.n_src > 1;
  const bool clamp_a = !G::is_static && p.concat_relu != 0;
  if (warp == 0) {
    // =============================== TMA producer: halo rows ===============================
    // one thread issues the per-row, per-K-block loads of local tile `it` into stage it % SA
    const int n_ks = G::is_static ? 1 : p.n_ks;            // halo K-slices per tile
    const int nkb_s = G::is_static ? g.nka() : p.nkb_s;    // halo K-blocks per slice (= all of them when n_ks == 1)
    auto issue_halo = [&](int si) {                        // si = it * n_ks + slice
      const int it = si / n_ks, kslice = si - it * n_ks;
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = si % p.SA;
      const int q0 = q_first + tile * kTileM;
      const int g_lo = (q0 - g.PH() * p.Wp - g.PW()) / p.Wp;
      const int g_hi = (q0 + kTileM - 1 + (g.KH() - 1 - g.PH()) * p.Wp + (g.KW() - 1 - g.PW())) / p.Wp;
      const int nrows = g_hi - g_lo + 1;
      const uint32_t full = smem_u32(&bar->a_full[s]);
      const uint32_t stage = sbase + p.off_a + s * p.a_stage_bytes;
      mbar_expect_tx(full, (uint32_t)(nrows * nkb_s * p.Wp * g.swa()));
      // row g >= 1 of the padded space is row hp = (g - 1) % Hp of image (g - 1) / Hp, i.e. source row h = hp - ZR
      // (negative: one of the zero rows above the image, filled by the TMA unit); g = 0 is all zero
      int n = (g_lo > 0) ? (g_lo - 1) / p.Hp : 0;
      int h = (g_lo > 0) ? (g_lo - 1) - n * p.Hp - p.ZR : -p.ZR - 1;  // -ZR - 1: the all-zero row above everything
      uint32_t dst = stage;
      const uint32_t row_bytes = p.Wp * g.swa();
      for (int r = 0; r < nrows; ++r, dst += row_bytes) {
        if (multi_src) {
          for (int kb = 0; kb < g.nka(); ++kb)
            tma_load_4d(dst + kb * p.a_kb_stride, &tmS.m[p.kb_src[kb]], full, (int)p.kb_c0[kb], 0, h, n);
        } else if (!G::is_static && p.n_box > 1) {
          for (int kb = 0; kb < nkb_s; ++kb)
            for (int bx = 0; bx < p.n_box; ++bx)
              tma_load_4d(dst + kb * p.a_kb_stride + bx * p.box_w * g.swa(), &tmS.m[0], full, (kslice * nkb_s + kb) * g.swa(), bx * p.box_w, h, n);
        } else {
#pragma unroll
          for (int kb = 0; kb < nkb_s; ++kb) tma_load_4d(dst + kb * p.a_kb_stride, &tmS.m[0], full, (kslice * nkb_s + kb) * g.swa(), 0, h, n);
        }
        if (h == -p.ZR - 1) {
          h = -p.ZR;  // g = 1: first row (zero row, if any) of image 0
        } else if (++h == p.H) {
          h = -p.ZR;  // zero rows shared by neighbouring images
          ++n;
        }
      }
    };
    if (!clamp_a) {
      if (elect_one()) {
        Tracer tr(p, 0);
        griddep_wait();  // the source may have been written by the previous kernel in the stream
        for (int si = 0; si < n_local * n_ks; ++si) {
          mbar_wait(smem_u32(&bar->a_empty[si % p.SA]), ((si / p.SA) & 1) ^ 1);
          tr.ev(1);
          issue_halo(si);
        }
      }
    } else {
      // Lane 0 issues the loads, the whole warp clamps: stage s of tile `it` is waited for (a_full), every byte
      // goes through a per-byte signed max with 0 (= vpmaxsb with zero, the reference's u8 / s8 concat ReLU),
      // the writes are made visible to the async proxy and a_ready[s] tells the MMA thread.  The next tile's
      // loads are issued BEFORE the clamp when their stage is already free, after it otherwise (the stage is
      // released by GEMM1 of tile it + 1 - SA, which itself may be waiting for this clamp).
      Tracer tr(p, 0);
      griddep_wait();
      if (lane == 0 && n_local > 0) issue_halo(0);  // stage 0 is free at kernel start
      for (int it = 0; it < n_local; ++it) {
        int issued = 1;
        if (it + 1 < n_local) {
          if (lane == 0) {
            issued = mbar_test_wait(smem_u32(&bar->a_empty[(it + 1) % p.SA]), (((it + 1) / p.SA) & 1) ^ 1) ? 1 : 0;
            if (issued) issue_halo(it + 1);
          }
          issued = __shfl_sync(0xffffffffu, issued, 0);
        }
        const int s = it % p.SA;
        mbar_wait_warp(smem_u32(&bar->a_full[s]), (it / p.SA) & 1);
        if (lane == 0) tr.ev(2);
        const uint32_t stage = sbase + p.off_a + s * p.a_stage_bytes;
        // (a_stage_bytes is a multiple of 1024: two 16-byte vectors per lane and step, four steps in flight.  The
        // clamp is two instructions per word: PRMT in sign-replicate mode builds 0xFF for every byte >= 128, LOP3
        // clears those bytes -- the producer warp is alone with this work, so its issue rate is what bounds it.)
#pragma unroll 4
        for (uint32_t off = (uint32_t)lane * 16u; off < p.a_stage_bytes; off += 1024u) {
          uint32_t v[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                         : "r"(stage + off + u * 512u));
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // == __vmaxs4(x, 0); (__byte_perm masks the selector's replicate bits: PTX)
              uint32_t neg;
              asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(neg) : "r"(v[u][i]));
              v[u][i] &= ~neg;
            }
            sts128(stage + off + u * 512u, v[u]);
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) {
          tr.ev(3);
          mbar_arrive(smem_u32(&bar->a_ready[s]));
          if (!issued) {
            mbar_wait(smem_u32(&bar->a_empty[(it + 1) % p.SA]), (((it + 1) / p.SA) & 1) ^ 1);
            issue_halo(it + 1);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // =============================== TMA producer: weights =================================
    if (elect_one()) {
      const bool c0_only = !G::is_static && p.conv0_only != 0;
      const int n_w0 = g.KH() * g.KW() * g.nkb(), n_w1 = c0_only ? 0 : g.n_chunks() * g.nkb1();
      if (g.w0_res() || g.w1_res()) {
        const uint32_t full = smem_u32(&bar->res_full);
        mbar_expect_tx(full, (g.w0_res() ? n_w0 * g.w0_block_bytes() : 0) + (g.w1_res() ? n_w1 * g.w1_block_bytes() : 0));
        if (g.w0_res())
          for (int b = 0; b < n_w0; ++b) tma_load_2d(sbase + p.off_w0 + b * g.w0_block_bytes(), &tmW0, full, 0, b * g.OC());
        if (g.w1_res())
          for (int b = 0; b < n_w1; ++b) tma_load_2d(sbase + p.off_w1 + b * g.w1_block_bytes(), &tmW1, full, 0, b * g.nc1());
      }
      if (!g.w0_res() || !g.w1_res()) {
        uint32_t s = 0, ph = 1;  // stage cursor and the parity to wait for on b_empty
        for (int it = 0; it <= n_local; ++it) {  // same interleaving as the MMA thread below
          if (it < n_local && !g.w0_res())
            for (int i = 0; i < n_w0; ++i) {
              int b = i;  // block id = tap * nkb + kb; with a K-sliced halo the MMA thread goes slice, tap, kb-in-slice
              if (!G::is_static && p.n_ks > 1) {
                const int per = n_w0 / p.n_ks, ksl = i / per, r = i - ksl * per, tap = r / p.nkb_s;
                b = tap * g.nkb() + ksl * p.nkb_s + (r - tap * p.nkb_s);
              }
              mbar_wait(smem_u32(&bar->b_empty[s]), ph);
              mbar_expect_tx(smem_u32(&bar->b_full[s]), g.w0_block_bytes());
              tma_load_2d(sbase + p.off_b + s * p.b_stage_bytes, &tmW0, smem_u32(&bar->b_full[s]), 0, b * g.OC());
              if (++s == (uint32_t)g.SB()) { s = 0; ph ^= 1; }
            }
          if (it >= 1 && !g.w1_res())
            for (int b = 0; b < n_w1; ++b) {
              mbar_wait(smem_u32(&bar->b_empty[s]), ph);
              mbar_expect_tx(smem_u32(&bar->b_full[s]), g.w1_block_bytes());
              tma_load_2d(sbase + p.off_b + s * p.b_stage_bytes, &tmW1, smem_u32(&bar->b_full[s]), 0, b * g.nc1());
              if (++s == (uint32_t)g.SB()) { s = 0; ph ^= 1; }
            }
        }
      }
      // ============================ GEMM2 issuer (all weights resident) ========================
      if (g.w0_res() && g.w1_res() && !c0_only) {
        const uint32_t idesc1 = make_idesc_i8(kTileM, g.nc1(), 0, 1);
        const uint64_t desc1_hi = make_smem_desc(0, 16, 8 * g.swb1(), layout_of(g.swb1()));
        const uint64_t w1_desc = desc1_hi | ((sbase + p.off_w1) >> 4);
        const uint64_t mid_desc = desc1_hi | ((sbase + p.off_mid) >> 4);
        const uint32_t w1_step = g.w1_block_bytes() >> 4, mid_step_kb = g.mid_kb_stride() >> 4;
        const int nks1_full = g.swb1() >> 5;
        mbar_wait(smem_u32(&bar->res_full), 0);
        uint32_t c = 0;
        for (int it = 0; it < n_local; ++it) {
          const int mb = it % p.NM;
          mbar_wait(smem_u32(&bar->mid_full[mb]), (it / p.NM) & 1);
          const uint64_t mid_it = mid_desc + (uint64_t)((mb * p.mid_bytes) >> 4);
          for (int j = 0; j < g.n_chunks(); ++j, ++c) {
            const uint32_t cb = c & 1;
            mbar_wait(smem_u32(&bar->acc1_empty[cb]), ((c >> 1) & 1) ^ (epilogue_seeds<G>() ? 0 : 1));  // seeded: handed over by the epilogue first
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem + kAcc1Col + cb * kAcc1Stride;
            if constexpr (seed_by_cp<G>()) seed_chunk_cp<false>(d_tmem, make_smem_desc(sbase + p.off_k1, 16, 128, kLayoutNone));
#pragma unroll
            for (int kb = 0; kb < g.nkb1(); ++kb) {
              const uint64_t b_desc = w1_desc + (uint64_t)((j * g.nkb1() + kb) * w1_step);
              const uint64_t a_desc = mid_it + kb * mid_step_kb;
              const int nks = (kb == g.nkb1() - 1) ? g.ks1_last() : nks1_full;
#pragma unroll
              for (int ks = 0; ks < nks; ++ks) if (!dbg_flag(p, 1)) umma_i8(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc1, seeded_acc1<G>() || (kb | ks) != 0);
            }
            umma_commit(smem_u32(&bar->acc1_full[cb]));
          }
          umma_commit(smem_u32(&bar->mid_empty[mb]));
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ======================================
    if (elect_one()) {
      const uint32_t idesc0 = make_idesc_i8(kTileM, g.OC(), 0, 1);
      const uint32_t idesc1 = make_idesc_i8(kTileM, g.nc1(), 0, 1);
      // descriptors differ only in their 14-bit start-address field (16 B units): constant part + adds
      const uint64_t desc0_hi = make_smem_desc(0, 16, 8 * g.swb(), layout_of(g.swb()));
      const uint64_t desc1_hi = make_smem_desc(0, 16, 8 * g.swb1(), layout_of(g.swb1()));
      const uint64_t desca_hi = make_smem_desc(0, 16, 8 * g.swa(), layout_of(g.swa()));  // A operand: its own K-block width
      const uint32_t a_step_kw = g.swa() >> 4, a_step_kh = (p.Wp * g.swa()) >> 4, a_step_kb = p.a_kb_stride >> 4;
      const int a_ks_per_block = g.swa() >> 5;  // 32-byte K-steps per A K-block
      const int a_ks_shift = a_ks_per_block == 4 ? 2 : (a_ks_per_block == 2 ? 1 : 0);
      const uint32_t w0_step = g.w0_block_bytes() >> 4, w1_step = g.w1_block_bytes() >> 4;
      const uint32_t mid_step_kb = g.mid_kb_stride() >> 4, b_stage_step = p.b_stage_bytes >> 4;
      const uint64_t w0_desc = desc0_hi | ((sbase + p.off_w0) >> 4), w1_desc = desc1_hi | ((sbase + p.off_w1) >> 4);
      const uint64_t bst0_desc = desc0_hi | ((sbase + p.off_b) >> 4), bst1_desc = desc1_hi | ((sbase + p.off_b) >> 4);
      const int nks_full = g.swb() >> 5, nks1_full = g.swb1() >> 5;
      if (g.w0_res() || g.w1_res()) mbar_wait(smem_u32(&bar->res_full), 0);
      // weight-stage cursor.  Static geometry: every GEMM starts at stage 0 (SB divides the number of
      // streamed blocks of each GEMM), so block i sits in stage i % SB -- a compile-time constant.
      uint32_t bs = 0, bph = 0;
      uint32_t c1count = 0;
      // position of the tile inside its halo window, advanced without divisions
      int a_off_px = (q_first - g.PH() * p.Wp - g.PW() + (int)blockIdx.x * kTileM) % p.Wp;
      uint32_t sa = 0, a_par = 0;
      // halo hand-off: the TMA's own barrier, or (fused concat + ReLU) the clamp's
      const uint32_t a_go = clamp_a ? smem_u32(&bar->a_ready[0]) : smem_u32(&bar->a_full[0]);
      Tracer tr(p, 1);
      tr.ev(9);

      // ---- taps [kw0, kw1) of tap row kh of GEMM1: K-blocks x K-steps each, descriptors by addition,
      //      weights in order
      uint32_t tap_i = 0;  // taps issued so far (GEMM1 throttle, all-resident plan only: see kG1Ahead)
      // (the conv-only operator has no conv1 chunks that could queue behind GEMM1: no throttle there)
      const bool throttle = g.w0_res() && g.w1_res() && !(!G::is_static && p.conv0_only != 0);
      const int n_ks = G::is_static ? 1 : p.n_ks, nkb_s = g.nkb() / n_ks;  // halo K-slices (Params::n_ks): weight K-blocks per slice
      auto gemm1_taps = [&](int kh, int kw0, int kw1, uint32_t d_tmem, uint64_t a_tile, int kslice = 0) {
        for (int kw = kw0; kw < kw1; ++kw) {
          if (throttle && tap_i >= (uint32_t)kG1Ahead)
            mbar_wait(smem_u32(&bar->g1_prog[tap_i % kG1Ahead]), ((tap_i / kG1Ahead) - 1) & 1);
#pragma unroll
          for (int kbl = 0; kbl < nkb_s; ++kbl) {
            const int kb = kslice * nkb_s + kbl;                // K-block of the reduction; kbl: its place in the halo stage
            const int blk = (kh * g.KW() + kw) * g.nkb() + kb;  // weight block index inside the tile
            const int first = (kh * g.KW() + kw) | kbl | kslice;  // 0 only for the tile's very first block
            uint64_t b_desc;
            uint32_t st = 0;
            if (g.w0_res()) {
              b_desc = w0_desc + blk * w0_step;
            } else {
              st = bs;
              mbar_wait(smem_u32(&bar->b_full[st]), bph);
              tc_fence_after_sync();
              b_desc = bst0_desc + st * b_stage_step;
            }
            const uint64_t a_tap = a_tile + kh * a_step_kh + kw * a_step_kw;
            const int nks = (kb == g.nkb() - 1) ? g.ks_last() : nks_full;
            if (G::is_static || g.swa() == g.swb()) {
              const uint64_t a_desc = a_tap + kbl * a_step_kb;
#pragma unroll
              for (int ks = 0; ks < nks; ++ks) if (!dbg_flag(p, 1)) umma_i8(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc0, (first | ks) != 0);
            } else {  // narrower A K-blocks (fused concat): K-step ks of weight block kb lives in A block k32 / a_ks_per_block
              // (swa divides swb, both powers of two: A block / step inside it by shift and mask, descriptors by addition)
              uint64_t a_blk = a_tap + (uint64_t)(((kb * nks_full) >> a_ks_shift) * a_step_kb);
              int kr = 0;
#pragma unroll 4
              for (int ks = 0; ks < nks; ++ks) {
                if (!dbg_flag(p, 1)) umma_i8(d_tmem, a_blk + 2 * kr, b_desc + 2 * ks, idesc0, (first | ks) != 0);
                if (++kr == a_ks_per_block) { kr = 0; a_blk += a_step_kb; }
              }
            }
            if (!g.w0_res()) {
              umma_commit(smem_u32(&bar->b_empty[st]));
              if (++bs == (uint32_t)g.SB()) { bs = 0; bph ^= 1; }
            }
          }
          if (throttle) {
            umma_commit(smem_u32(&bar->g1_prog[tap_i % kG1Ahead]));
            ++tap_i;
          }
        }
      };
      // ---- one N-chunk of GEMM2 with W1 either resident or from the ring
      auto gemm2_chunk = [&](int j, uint32_t d_tmem, uint64_t mid_desc) {
        if constexpr (seed_by_cp<G>()) seed_chunk_cp<false>(d_tmem, make_smem_desc(sbase + p.off_k1, 16, 128, kLayoutNone));
#pragma unroll
        for (int kb = 0; kb < g.nkb1(); ++kb) {
          const int blk = j * g.nkb1() + kb;
          uint64_t b_desc;
          uint32_t st = 0;
          if (g.w1_res()) {
            b_desc = w1_desc + blk * w1_step;
          } else {
            st = bs;
            mbar_wait(smem_u32(&bar->b_full[st]), bph);
            tc_fence_after_sync();
            b_desc = bst1_desc + st * b_stage_step;
          }
          const uint64_t a_desc = mid_desc + kb * mid_step_kb;
          const int nks = (kb == g.nkb1() - 1) ? g.ks1_last() : nks1_full;
#pragma unroll
          for (int ks = 0; ks < nks; ++ks) if (!dbg_flag(p, 1)) umma_i8(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc1, seeded_acc1<G>() || (kb | ks) != 0);
          if (!g.w1_res()) {
            umma_commit(smem_u32(&bar->b_empty[st]));
            if (++bs == (uint32_t)g.SB()) { bs = 0; bph ^= 1; }
          }
        }
      };

      if ((g.w1_res() && g.w0_res()) || (!G::is_static && p.conv0_only)) {
        // (conv0-only operator: there is no GEMM2 at all.)
        // All weights resident: the two GEMM streams are independent and each has its own issuing
        // thread: this one runs GEMM1 as far ahead as accumulators and halo stages allow, warp 2 (done
        // with the weight loads) issues every conv1 chunk the moment its accumulator is free.  One
        // thread polling for both spent ~110 cycles per MMA on issue and left the epilogue waiting behind
        // whatever GEMM1 work was queued (profiles/r01_trace_cfg3_v8.log).
        for (int it = 0; it < n_local; ++it) {
          const int ab = it % g.n_acc0();
          mbar_wait(smem_u32(&bar->acc0_empty[ab]), ((it / g.n_acc0()) & 1) ^ 1);
          const uint32_t d_tmem = tmem + ab * g.OC();
          for (int kslice = 0; kslice < n_ks; ++kslice) {  // (one pass unless the halo is K-sliced)
            mbar_wait(a_go + 8 * sa, a_par);
            tc_fence_after_sync();
            tr.ev(10);
            const uint64_t a_tile = desca_hi | ((sbase + p.off_a + sa * p.a_stage_bytes + a_off_px * g.swa()) >> 4);
#pragma unroll
            for (int kh = 0; kh < g.KH(); ++kh) gemm1_taps(kh, 0, g.KW(), d_tmem, a_tile, kslice);
            umma_commit(smem_u32(&bar->a_empty[sa]));
            if (++sa == (uint32_t)p.SA) { sa = 0; a_par ^= 1; }
          }
          umma_commit(smem_u32(&bar->acc0_full[ab]));
          tr.ev(11);
          a_off_px += p.tile_step_mod;
          if (a_off_px >= p.Wp) a_off_px -= p.Wp;
        }
      } else if (g.w1_res() && !g.w0_res()) {
        // W1 resident, W0 streamed by warp 2: one thread issues both GEMMs, readiness-driven -- a GEMM2
        // chunk whenever its accumulator is free and the intermediate tile is there (it unblocks the
        // epilogue, the longer path), otherwise the next tap row of the next GEMM1.  A fixed order
        // suffers head-of-line blocking in both directions (profiles/r01_trace_cfg3_v4.log).
        int g1_it = 0, g1_kh = 0, g2_it = 0, g2_j = 0;
        bool g2_open = false;
        uint64_t a_tile = 0;
        uint32_t d0 = 0;
        uint32_t idle = 0;
        while (g2_it < n_local) {
          bool did = false;
          if (++idle > (1u << 27)) {
            printf("conv MMA scheduler stuck: block %d g1 %d/%d g2 %d/%d\n", blockIdx.x, g1_it, g1_kh, g2_it, g2_j);
            __trap();
          }
          if (g2_it < g1_it) {  // GEMM1(g2_it) has been issued completely
            const int mb = g2_it % p.NM;
            if (!g2_open && mbar_test_wait(smem_u32(&bar->mid_full[mb]), (g2_it / p.NM) & 1)) {
              g2_open = true;
              tr.ev(12);
            }
            if (g2_open) {
              const int cb = c1count & 1;
              if (mbar_test_wait(smem_u32(&bar->acc1_empty[cb]), ((c1count >> 1) & 1) ^ (epilogue_seeds<G>() ? 0 : 1))) {
                tc_fence_after_sync();
                const uint64_t mid_desc = desc1_hi | ((sbase + p.off_mid + mb * p.mid_bytes) >> 4);
                gemm2_chunk(g2_j, tmem + kAcc1Col + cb * kAcc1Stride, mid_desc);
                umma_commit(smem_u32(&bar->acc1_full[cb]));
                tr.ev(13);
                ++c1count;
                did = true;
                idle = 0;
                if (++g2_j == g.n_chunks()) {
                  umma_commit(smem_u32(&bar->mid_empty[mb]));
                  g2_j = 0;
                  g2_open = false;
                  ++g2_it;
                }
              }
            }
          }
          if (!did && g1_it < n_local) {
            bool ok = true;
            if (g1_kh == 0) {
              const int ab = g1_it % g.n_acc0();
              ok = mbar_test_wait(smem_u32(&bar->acc0_empty[ab]), ((g1_it / g.n_acc0()) & 1) ^ 1) &&
                   mbar_test_wait(a_go + 8 * sa, a_par);
              if (ok) {
                tc_fence_after_sync();
                tr.ev(10);
                d0 = tmem + ab * g.OC();
                a_tile = desca_hi | ((sbase + p.off_a + sa * p.a_stage_bytes + a_off_px * g.swa()) >> 4);
              }
            }
            if (ok) {
              gemm1_taps(g1_kh, 0, g.KW(), d0, a_tile);
              did = true;
              idle = 0;
              if (++g1_kh == g.KH()) {
                umma_commit(smem_u32(&bar->a_empty[sa]));
                umma_commit(smem_u32(&bar->acc0_full[g1_it % g.n_acc0()]));
                tr.ev(11);
                g1_kh = 0;
                ++g1_it;
                if (++sa == (uint32_t)p.SA) { sa = 0; a_par ^= 1; }
                a_off_px += p.tile_step_mod;
                if (a_off_px >= p.Wp) a_off_px -= p.Wp;
              }
            }
          }
        }
      } else {
        // W0 and W1 share one ring: the order is fixed -- GEMM1(it) then GEMM2(it-1)
        for (int it = 0; it <= n_local; ++it) {
          if (it < n_local) {
            const int ab = it % g.n_acc0();
            mbar_wait(smem_u32(&bar->acc0_empty[ab]), ((it / g.n_acc0()) & 1) ^ 1);
            const uint32_t d_tmem = tmem + ab * g.OC();
            for (int kslice = 0; kslice < n_ks; ++kslice) {  // (one pass unless the halo is K-sliced)
              mbar_wait(a_go + 8 * sa, a_par);
              tc_fence_after_sync();
              tr.ev(10);
              const uint64_t a_tile = desca_hi | ((sbase + p.off_a + sa * p.a_stage_bytes + a_off_px * g.swa()) >> 4);
#pragma unroll
              for (int kh = 0; kh < g.KH(); ++kh) gemm1_taps(kh, 0, g.KW(), d_tmem, a_tile, kslice);
              umma_commit(smem_u32(&bar->a_empty[sa]));
              if (++sa == (uint32_t)p.SA) { sa = 0; a_par ^= 1; }
            }
            umma_commit(smem_u32(&bar->acc0_full[ab]));
            tr.ev(11);
            a_off_px += p.tile_step_mod;
            if (a_off_px >= p.Wp) a_off_px -= p.Wp;
          }
          if (it >= 1) {
            const int jt = it - 1, mb = jt % p.NM;
            mbar_wait(smem_u32(&bar->mid_full[mb]), (jt / p.NM) & 1);
            tc_fence_after_sync();
            tr.ev(12);
            const uint64_t mid_desc = desc1_hi | ((sbase + p.off_mid + mb * p.mid_bytes) >> 4);
            for (int j = 0; j < g.n_chunks(); ++j, ++c1count) {
              const int cb = c1count & 1;
              mbar_wait(smem_u32(&bar->acc1_empty[cb]), ((c1count >> 1) & 1) ^ (epilogue_seeds<G>() ? 0 : 1));
              tc_fence_after_sync();
              gemm2_chunk(j, tmem + kAcc1Col + cb * kAcc1Stride, mid_desc);
              umma_commit(smem_u32(&bar->acc1_full[cb]));
              tr.ev(13);
            }
            umma_commit(smem_u32(&bar->mid_empty[mb]));
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ====================================== epilogue =======================================
    load_epilogue_constants<G>(p, smem);
    if constexpr (static_two_groups<G>())
      epilogue_static2<G, kDst, false>(p, smem, bar, tmem, warp, lane, n_local, (int)blockIdx.x, (int)gridDim.x);
    else if constexpr (static_epilogue<G>())
      epilogue_static<G, kDst, false>(p, tmD, smem, bar, tmem, warp, lane, n_local, (int)blockIdx.x, (int)gridDim.x);
    else
      epilogue_role<G, kDst, kDown0, kDown1, kNanSafe, false>(p, smem, bar, tmem, warp, lane, n_local, (int)blockIdx.x,
                                                              (int)gridDim.x);
  } else if (warp == 3) {
    // ====================== store thread (staged output, run-time geometry only) ================
    if constexpr (!G::is_static) {
      if ((kDst == DF_U8 || kDst == DF_S8) && p.stage_out && elect_one())
        store_role<G>(p, tmD, bar, sbase, n_local, (int)blockIdx.x, (int)gridDim.x);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 3) tmem_dealloc<512>(tmem);
  trace_wallclock(p, 1);
}

// =============================================================================== CTA-pair kernel
// cta_group::2 version for shapes whose weights fit in shared memory once they are SPLIT across a
// pair of CTAs (BASELINE cfg3: 144 KB + 64 KB -> 72 + 32 KB per CTA): no weight streaming at all.
// A cluster of 2 CTAs works on two adjacent 128-position tiles as one M = 256 MMA tile:
//   * each CTA loads its own halo (TMA, completion bytes reported to the LEADER's barrier), owns the
//     128 TMEM lanes of its tile and runs its own 16 epilogue warps;
//   * each CTA holds rows [64r, 64r+64) of every weight block; the tensor cores of both SMs read both
//     halves (that is what cta_group::2 does), so B is fetched from HBM/L2 once per pair and never again;
//   * only the leader issues tcgen05.mma / tcgen05.commit (multicast to both CTAs' barriers); the
//     peer's epilogue warps arrive on the leader's barriers through shared::cluster addresses.
// One descriptor must address both CTAs' halo buffers, so instead of shifting the descriptor by the
// tile's offset inside its halo window, every CTA shifts the TMA DESTINATION such that its tile origin
// always lands at the same shared-memory offset.
struct PairBarriers {
  uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  uint64_t res_full[4], peer_ready[4];  // resident weights arrive in four parts: W0 tap rows 0..2, W1
  uint64_t b_full[kMaxBStages], b_empty[kMaxBStages];  // streamed weight halves (G::w0_res == 0): ring of stages
  uint64_t acc0_full[2], acc0_empty[2];
  uint64_t mid_full[2], mid_empty[2];
  uint64_t acc1_full[2], acc1_empty[2];
  uint64_t stage_full[2], stage_empty[2];
  uint64_t g1_prog[kG1Ahead];  // GEMM1 issue throttle (see G1Throttle)
  uint32_t tmem_base;
};

template <class G, int kDst>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ DstMaps tmD,
                 const __grid_constant__ Params p) {
  // Static geometries: BASELINE cfg3 (resident) / cfg4 (streamed).  Run-time geometry (DynGeom): any shape of the
  // single-CTA kernel's stride-1-window family whose weights have to stream -- always the streamed form, generic
  // epilogue (epilogue_role), taps / K-blocks / chunk counts from Params.
  // G::w0_res == 1: the weight halves are resident for the whole kernel (BASELINE cfg3).
  // G::w0_res == 0: they stream through a ring of stages (BASELINE cfg4: 576 + 256 KB of weights).  Each CTA loads
  //   ITS half of every block (rows [r * N/2, (r+1) * N/2)), so a pair pulls every weight byte out of the L2 once
  //   per 256 positions instead of once per 128 -- the single-CTA kernel is bound by exactly that traffic -- and a
  //   stage of S bytes feeds twice the MMA work, so the latency-bound ring delivers twice the rate.  One thread of
  //   the leader issues both GEMMs in the ring's fixed order: GEMM1(it), then GEMM2(it - 1).
  constexpr bool kResident = [] {
    if constexpr (G::is_static) return G::w0_res != 0;
    else return false;
  }();
  const Geo<G> g{p};
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  PairBarriers* bar = reinterpret_cast<PairBarriers*>(smem);
  const uint32_t sbase = smem_u32(smem);
  trace_wallclock(p, 0);
  griddep_launch_dependents();  // PDL, see conv_fused_kernel
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int n_pair_tiles = (p.n_tiles + 1) >> 1;
  const int n_local = (n_pair_tiles - cid + ncl - 1) / ncl;
  const int kHalfRows0 = g.OC() / 2, kHalfRows1 = g.nc1() / 2;
  const uint32_t kW0Half = kHalfRows0 * g.swb(), kW1Half = kHalfRows1 * g.swb1();
  const bool c0_only = !G::is_static && p.conv0_only != 0;  // conv-only operator: no GEMM2 at all
  const int kNW0 = g.KH() * g.KW() * g.nkb(), kNW1 = c0_only ? 0 : g.n_chunks() * g.nkb1();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(smem_u32(&bar->a_full[i]), 2);  // one expect_tx arrival per CTA of the pair
      mbar_init(smem_u32(&bar->a_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bar->res_full[i]), 1);
      mbar_init(smem_u32(&bar->peer_ready[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar->acc0_full[i]), 1);
      mbar_init(smem_u32(&bar->acc0_empty[i]), 2 * unit_warps<G>());
      mbar_init(smem_u32(&bar->mid_full[i]), 2 * unit_warps<G>());
      mbar_init(smem_u32(&bar->mid_empty[i]), 1);
      mbar_init(smem_u32(&bar->acc1_full[i]), 1);
      mbar_init(smem_u32(&bar->acc1_empty[i]), 2 * unit_warps<G>());
      mbar_init(smem_u32(&bar->stage_full[i]), kUnitWarps);
      mbar_init(smem_u32(&bar->stage_empty[i]), 1);
    }
    for (int i = 0; i < kG1Ahead; ++i) mbar_init(smem_u32(&bar->g1_prog[i]), 1);
    for (int i = 0; i < kMaxBStages; ++i) {
      mbar_init(smem_u32(&bar->b_full[i]), 2);   // one expect_tx arrival per CTA of the pair (leader's copy is used)
      mbar_init(smem_u32(&bar->b_empty[i]), 1);  // multicast commit of the MMAs that read the stage
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW0);
    tma_prefetch_desc(&tmW1);
  }
  if (warp == 3) tmem_alloc_pair<512>(smem_u32(&bar->tmem_base));
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anyone arrives remotely
  tc_fence_after_sync();
  const uint32_t tmem = bar->tmem_base;
  const int q_first = p.q_first;
  const uint32_t a_origin = (uint32_t)p.Wp * g.swb();  // tile origin inside a halo stage (same in both CTAs)

  if (warp == 0) {
    // ================================ halo producer (both CTAs) ================================
    if (elect_one()) {
      Tracer tr(p, 0);
      griddep_wait();  // the source may have been written by the previous kernel in the stream
      for (int it = 0; it < n_local; ++it) {
        const int tile = 2 * (cid + it * ncl) + (int)rank;
        const int s = it % p.SA;
        mbar_wait(smem_u32(&bar->a_empty[s]), ((it / p.SA) & 1) ^ 1);
        tr.ev(1);
        const int q0 = q_first + tile * kTileM;
        const int q_halo = q0 - g.PH() * p.Wp - g.PW();  // first position the tile's taps read
        const int g_lo = q_halo / p.Wp;
        const int g_hi = (q0 + kTileM - 1 + (g.KH() - 1 - g.PH()) * p.Wp + (g.KW() - 1 - g.PW())) / p.Wp;
        const int nrows = g_hi - g_lo + 1;
        const int a_off_px = q_halo - g_lo * p.Wp;
        const uint32_t leader_full = mapa_u32(smem_u32(&bar->a_full[s]), 0);
        mbar_expect_tx_cluster(leader_full, (uint32_t)(nrows * g.nkb() * p.Wp * g.swb()));
        int n = (g_lo > 0) ? (g_lo - 1) / p.Hp : 0;
        int h = (g_lo > 0) ? (g_lo - 1) - n * p.Hp - p.ZR : -p.ZR - 1;  // see conv_fused_kernel's halo producer
        uint32_t dst = sbase + p.off_a + s * p.a_stage_bytes + a_origin - (uint32_t)a_off_px * g.swb();
        const uint32_t row_bytes = p.Wp * g.swb();
        for (int r = 0; r < nrows; ++r, dst += row_bytes) {
#pragma unroll
          for (int kb = 0; kb < g.nkb(); ++kb) tma_load_4d_pair(dst + kb * p.a_kb_stride, &tmA, leader_full, kb * g.swb(), 0, h, n);
          if (h == -p.ZR - 1) {
            h = -p.ZR;
          } else if (++h == p.H) {
            h = -p.ZR;
            ++n;
          }
        }
      }
    }
  } else if (warp == 2) {
   if constexpr (!kResident) {
    // ============================ streamed weight halves (both CTAs) ===========================
    if (elect_one()) {
      uint32_t s = 0, ph = 1;  // stage cursor and the parity to wait for on this CTA's b_empty
      auto load_half = [&](const CUtensorMap* tm, uint32_t bytes, int row) __attribute__((always_inline)) {
        mbar_wait(smem_u32(&bar->b_empty[s]), ph);
        const uint32_t leader_full = mapa_u32(smem_u32(&bar->b_full[s]), 0);
        mbar_expect_tx_cluster(leader_full, bytes);
        tma_load_2d_pair(sbase + p.off_b + s * p.b_stage_bytes, tm, leader_full, 0, row);
        if (++s == (uint32_t)p.SB) { s = 0; ph ^= 1; }
      };
      auto load_w0_tap = [&](int t) __attribute__((always_inline)) {
        for (int kb = 0; kb < g.nkb(); ++kb) load_half(&tmW0, kW0Half, (t * g.nkb() + kb) * g.OC() + (int)rank * kHalfRows0);
      };
      // conv1 half-blocks are smaller than a stage (half as big for the BASELINE shapes): w1_pack of them share one, so
      // that a stage feeds as many MMA cycles in GEMM2 as in GEMM1 and the ring's refill latency is covered there too
      auto load_w1_chunk = [&](int j) __attribute__((always_inline)) {
        const int pack = p.w1_pack;
        for (int b = j * g.nkb1(); b < (j + 1) * g.nkb1(); b += pack) {
          mbar_wait(smem_u32(&bar->b_empty[s]), ph);
          const uint32_t leader_full = mapa_u32(smem_u32(&bar->b_full[s]), 0);
          mbar_expect_tx_cluster(leader_full, (uint32_t)pack * kW1Half);
          for (int q = 0; q < pack; ++q)
            tma_load_2d_pair(sbase + p.off_b + s * p.b_stage_bytes + q * kW1Half, &tmW1, leader_full, 0, (b + q) * g.nc1() + (int)rank * kHalfRows1);
          if (++s == (uint32_t)p.SB) { s = 0; ph ^= 1; }
        }
      };
      const int n_taps = g.KH() * g.KW(), n_ch = c0_only ? 0 : g.n_chunks();
      for (int it = 0; it <= n_local; ++it) {  // same order as the MMA thread below (see there)
        const bool has_g1 = it < n_local, has_g2 = it >= 1 && n_ch > 0;
        if (p.g_interleave && has_g1 && has_g2) {
          load_w0_tap(0);
          int t = 1;
          for (int j = 0; j < n_ch; ++j) {
            load_w1_chunk(j);
            for (const int upto = 1 + ((j + 1) * (n_taps - 1)) / n_ch; t < upto; ++t) load_w0_tap(t);
          }
        } else {
          if (has_g1)
            for (int t = 0; t < n_taps; ++t) load_w0_tap(t);
          if (has_g2)
            for (int j = 0; j < n_ch; ++j) load_w1_chunk(j);
        }
      }
    }
   } else {
    // ============================ resident weight halves (both CTAs) ===========================
    if (elect_one()) {
      // four parts with their own barriers, so that GEMM1 of the first tile starts after a third of W0
      constexpr int kRowBlocks = 3 * G::nkb;
      for (int part = 0; part < 3; ++part) {
        const uint32_t full = smem_u32(&bar->res_full[part]);
        mbar_expect_tx(full, kRowBlocks * kW0Half);
        for (int b = part * kRowBlocks; b < (part + 1) * kRowBlocks; ++b)
          tma_load_2d(sbase + p.off_w0 + b * kW0Half, &tmW0, full, 0, b * G::OC + (int)rank * kHalfRows0);
      }
      {
        const uint32_t full = smem_u32(&bar->res_full[3]);
        mbar_expect_tx(full, kNW1 * kW1Half);
        for (int b = 0; b < kNW1; ++b) tma_load_2d(sbase + p.off_w1 + b * kW1Half, &tmW1, full, 0, b * G::nc1 + (int)rank * kHalfRows1);
      }
      for (int part = 0; part < 4; ++part) {
        mbar_wait(smem_u32(&bar->res_full[part]), 0);
        if (rank == 1) mbar_arrive_cluster(mapa_u32(smem_u32(&bar->peer_ready[part]), 0));
      }
      // ============================== GEMM2 issuer (leader only) ==============================
      // Its own thread, so that a conv1 chunk -- the thing the epilogue is waiting for -- is issued the
      // moment its accumulator is free, instead of after whatever GEMM1 work a single scheduler thread
      // happens to be in the middle of (a polling scheduler spent ~110 cycles per MMA on issue and left
      // the tensor pipe idle: profiles/r01_trace_cfg3_v8.log).
      if (rank == 0) {
        const uint32_t idesc1 = make_idesc_i8(2 * kTileM, G::nc1, 0, 1);
        const uint64_t desc1_hi = make_smem_desc(0, 16, 8 * G::swb1, layout_of(G::swb1));
        const uint64_t w1_desc = desc1_hi | ((sbase + p.off_w1) >> 4);
        const uint64_t mid_desc = desc1_hi | ((sbase + p.off_mid) >> 4);
        Tracer tr(p, 1);
        mbar_wait(smem_u32(&bar->peer_ready[3]), 0);
        uint32_t c = 0;
        for (int it = 0; it < n_local; ++it) {
          const int mb = it % p.NM;
          mbar_wait(smem_u32(&bar->mid_full[mb]), (it / p.NM) & 1);
          const uint64_t mid_it = mid_desc + (uint64_t)((mb * p.mid_bytes) >> 4);
          tr.ev(12);
#pragma unroll
          for (int j = 0; j < G::n_chunks; ++j, ++c) {
            const uint32_t cb = c & 1;
            mbar_wait(smem_u32(&bar->acc1_empty[cb]), ((c >> 1) & 1) ^ (epilogue_seeds<G>() ? 0 : 1));  // seeded: handed over by the epilogue first
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem + kAcc1Col + cb * kAcc1Stride;
            if constexpr (seed_by_cp<G>()) seed_chunk_cp<true>(d_tmem, make_smem_desc(sbase + p.off_k1, 16, 128, kLayoutNone));
#pragma unroll
            for (int kb = 0; kb < G::nkb1; ++kb) {
              const uint64_t b_desc = w1_desc + (uint64_t)((j * G::nkb1 + kb) * (kW1Half >> 4));
              const uint64_t a_desc = mid_it + kb * ((kTileM * G::swb1) >> 4);
              constexpr int nks_full = G::swb1 >> 5;
              const int nks = (kb == G::nkb1 - 1) ? G::ks1_last : nks_full;
#pragma unroll
              for (int ks = 0; ks < nks; ++ks)
                if (!dbg_flag(p, 1)) umma_i8_pair(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc1, seeded_acc1<G>() || (kb | ks) != 0);
            }
            umma_commit_pair(smem_u32(&bar->acc1_full[cb]));
            tr.ev(13);
          }
          umma_commit_pair(smem_u32(&bar->mid_empty[mb]));
        }
      }
    }
   }
  } else if (warp == 1) {
   if constexpr (!kResident) {
    // ====================== MMA issuer, streamed weights (leader only): GEMM1(it), GEMM2(it - 1) ======================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc0 = make_idesc_i8(2 * kTileM, g.OC(), 0, 1), idesc1 = make_idesc_i8(2 * kTileM, g.nc1(), 0, 1);
      const uint64_t desc0_hi = make_smem_desc(0, 16, 8 * g.swb(), layout_of(g.swb()));
      const uint64_t desc1_hi = make_smem_desc(0, 16, 8 * g.swb1(), layout_of(g.swb1()));
      const uint32_t a_step_kw = g.swb() >> 4, a_step_kh = (p.Wp * g.swb()) >> 4, a_step_kb = p.a_kb_stride >> 4;
      const uint64_t bst0_desc = desc0_hi | ((sbase + p.off_b) >> 4), bst1_desc = desc1_hi | ((sbase + p.off_b) >> 4);
      const uint64_t mid_desc = desc1_hi | ((sbase + p.off_mid) >> 4);
      const uint32_t b_stage_step = p.b_stage_bytes >> 4;
      Tracer tr(p, 1);
      tr.ev(9);
      uint32_t sa = 0, a_par = 0, bs = 0, bph = 0, c1count = 0;
      // GEMM1 of tile it: begin (accumulator + halo ready), one tap at a time, end (halo stage and accumulator handed on)
      int g1_ab = 0;
      uint32_t g1_d0 = 0;
      uint64_t g1_a_tile = 0;
      auto g1_begin = [&](int it) __attribute__((always_inline)) {
        g1_ab = it % g.n_acc0();
        mbar_wait(smem_u32(&bar->acc0_empty[g1_ab]), ((it / g.n_acc0()) & 1) ^ 1);
        mbar_wait(smem_u32(&bar->a_full[sa]), a_par);
        tc_fence_after_sync();
        tr.ev(10);
        g1_d0 = tmem + g1_ab * g.OC();
        g1_a_tile = desc0_hi | ((sbase + p.off_a + sa * p.a_stage_bytes + a_origin) >> 4);
      };
      auto g1_tap = [&](int kh, int kw) __attribute__((always_inline)) {
#pragma unroll
        for (int kb = 0; kb < g.nkb(); ++kb) {
          const int blk = (kh * g.KW() + kw) * g.nkb() + kb;
          mbar_wait(smem_u32(&bar->b_full[bs]), bph);
          tc_fence_after_sync();
          const uint64_t b_desc = bst0_desc + (uint64_t)(bs * b_stage_step);
          const uint64_t a_desc = g1_a_tile + kh * a_step_kh + kw * a_step_kw + kb * a_step_kb;
          const int nks_full = g.swb() >> 5;
          const int nks = (kb == g.nkb() - 1) ? g.ks_last() : nks_full;
#pragma unroll
          for (int ks = 0; ks < nks; ++ks)
            if (!dbg_flag(p, 1)) umma_i8_pair(g1_d0, a_desc + 2 * ks, b_desc + 2 * ks, idesc0, (blk | ks) != 0);
          umma_commit_pair(smem_u32(&bar->b_empty[bs]));
          if (++bs == (uint32_t)p.SB) { bs = 0; bph ^= 1; }
        }
      };
      auto g1_end = [&]() __attribute__((always_inline)) {
        umma_commit_pair(smem_u32(&bar->a_empty[sa]));
        umma_commit_pair(smem_u32(&bar->acc0_full[g1_ab]));
        tr.ev(11);
        if (++sa == (uint32_t)p.SA) { sa = 0; a_par ^= 1; }
      };
      // GEMM2 of tile jt: begin (intermediate tile ready), one chunk at a time, end (intermediate tile handed back)
      int g2_mb = 0;
      uint64_t g2_mid = 0;
      auto g2_begin = [&](int jt) __attribute__((always_inline)) {
        g2_mb = jt % p.NM;
        mbar_wait(smem_u32(&bar->mid_full[g2_mb]), (jt / p.NM) & 1);
        tc_fence_after_sync();
        tr.ev(12);
        g2_mid = mid_desc + (uint64_t)((g2_mb * p.mid_bytes) >> 4);
      };
      auto g2_chunk = [&]() __attribute__((always_inline)) {
        const uint32_t cb = c1count & 1;
        mbar_wait(smem_u32(&bar->acc1_empty[cb]), ((c1count >> 1) & 1) ^ (epilogue_seeds<G>() ? 0 : 1));
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem + kAcc1Col + cb * kAcc1Stride;
        if constexpr (seed_by_cp<G>()) seed_chunk_cp<true>(d_tmem, make_smem_desc(sbase + p.off_k1, 16, 128, kLayoutNone));
        const int pack = p.w1_pack;
#pragma unroll
        for (int kb = 0; kb < g.nkb1(); ++kb) {
          const int q = kb & (pack - 1);  // position inside the ring stage (pack is 1 or 2, see the loader)
          if (q == 0) {
            mbar_wait(smem_u32(&bar->b_full[bs]), bph);
            tc_fence_after_sync();
          }
          const uint64_t b_desc = bst1_desc + (uint64_t)(bs * b_stage_step) + (uint64_t)((q * kW1Half) >> 4);
          const uint64_t a_desc = g2_mid + kb * ((kTileM * g.swb1()) >> 4);
          const int nks_full = g.swb1() >> 5;
          const int nks = (kb == g.nkb1() - 1) ? g.ks1_last() : nks_full;
#pragma unroll
          for (int ks = 0; ks < nks; ++ks)
            if (!dbg_flag(p, 1)) umma_i8_pair(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc1, seeded_acc1<G>() || (kb | ks) != 0);
          if (q == pack - 1) {
            umma_commit_pair(smem_u32(&bar->b_empty[bs]));
            if (++bs == (uint32_t)p.SB) { bs = 0; bph ^= 1; }
          }
        }
        umma_commit_pair(smem_u32(&bar->acc1_full[cb]));
        tr.ev(13);
        ++c1count;
      };
      auto g2_end = [&]() __attribute__((always_inline)) { umma_commit_pair(smem_u32(&bar->mid_empty[g2_mb])); };
      // Order of issue = order of the ring.  GEMM1(it) as a block followed by GEMM2(it - 1) as a block serialises the two
      // pipelines: during GEMM1 the epilogue has nothing but E0 to do, during GEMM2 the tensor pipe waits for the
      // epilogue chunk by chunk (two conv1 accumulators) -- a cfg4 pair-tile took 22 k cycles for 13.3 k of MMA and ~11 k
      // of epilogue.  Interleaved (g_interleave), chunk j of GEMM2(it - 1) is followed by the next ~taps / chunks taps of
      // GEMM1(it): while the epilogue works on a chunk the tensor pipe runs a tap.  Dependencies are unchanged: the first
      // tap needs E0(it - 1) to have read the conv0 accumulator (it has: chunk 0 needed its result), E0(it) needs the
      // last tap, which follows the last chunk, which needs C_{n-3}(it - 1) -- all earlier in the epilogue's unit stream.
      const int n_taps = g.KH() * g.KW(), n_ch = c0_only ? 0 : g.n_chunks();
      for (int it = 0; it <= n_local; ++it) {
        const bool has_g1 = it < n_local, has_g2 = it >= 1 && n_ch > 0;
        if (p.g_interleave && has_g1 && has_g2) {
          int t = 0, kh = 0, kw = 0;
          auto taps_upto = [&](int upto) __attribute__((always_inline)) {
            for (; t < upto; ++t) {
              g1_tap(kh, kw);
              if (++kw == g.KW()) { kw = 0; ++kh; }
            }
          };
          g1_begin(it);  // the first tap goes ahead of chunk 0: it only needs E0(it - 1) to have READ the accumulator
          taps_upto(1);
          g2_begin(it - 1);
          for (int j = 0; j < n_ch; ++j) {
            g2_chunk();
            taps_upto(1 + ((j + 1) * (n_taps - 1)) / n_ch);
          }
          g2_end();
          g1_end();
        } else {
          if (has_g1) {
            g1_begin(it);
#pragma unroll
            for (int kh = 0; kh < g.KH(); ++kh) {
#pragma unroll
              for (int kw = 0; kw < g.KW(); ++kw) g1_tap(kh, kw);
            }
            g1_end();
          }
          if (has_g2) {
            g2_begin(it - 1);
            for (int j = 0; j < n_ch; ++j) g2_chunk();
            g2_end();
          }
        }
      }
    }
   } else {
    // ================================ GEMM1 issuer (leader only) ================================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc0 = make_idesc_i8(2 * kTileM, G::OC, 0, 1);
      const uint64_t desc0_hi = make_smem_desc(0, 16, 8 * G::swb, layout_of(G::swb));
      const uint32_t a_step_kw = G::swb >> 4, a_step_kh = (p.Wp * G::swb) >> 4, a_step_kb = p.a_kb_stride >> 4;
      const uint64_t w0_desc = desc0_hi | ((sbase + p.off_w0) >> 4);
      Tracer tr(p, 1);
      tr.ev(9);
      uint32_t sa = 0, a_par = 0, tap_i = 0;
      for (int it = 0; it < n_local; ++it) {
        const int ab = it & 1;
        mbar_wait(smem_u32(&bar->acc0_empty[ab]), ((it >> 1) & 1) ^ 1);
        mbar_wait(smem_u32(&bar->a_full[sa]), a_par);
        tc_fence_after_sync();
        tr.ev(10);
        const uint32_t d0 = tmem + ab * G::OC;
        const uint64_t a_tile = desc0_hi | ((sbase + p.off_a + sa * p.a_stage_bytes + a_origin) >> 4);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if (it == 0) {  // first tile: this tap row's weights must have landed in both CTAs
            mbar_wait(smem_u32(&bar->res_full[kh]), 0);
            mbar_wait(smem_u32(&bar->peer_ready[kh]), 0);
            tc_fence_after_sync();
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            // throttle, see kG1Ahead.  Not for the first tile: no conv1 chunk can be waiting behind it, and the waits
            // stretch its GEMM1 -- the whole pipeline's fill -- from 2304 to ~3300 cycles.  (Skipped waits leave
            // the two progress barriers a phase behind for a tap or two; the throttle is advisory, not a dependency.)
            if (DF_G1_FREE_FIRST ? (it > 0 && tap_i >= (uint32_t)kG1Ahead) : (tap_i >= (uint32_t)kG1Ahead))
              mbar_wait(smem_u32(&bar->g1_prog[tap_i % kG1Ahead]), ((tap_i / kG1Ahead) - 1) & 1);
#pragma unroll
            for (int kb = 0; kb < G::nkb; ++kb) {
              const int blk = (kh * 3 + kw) * G::nkb + kb;
              const uint64_t b_desc = w0_desc + (uint64_t)(blk * (kW0Half >> 4));
              const uint64_t a_desc = a_tile + kh * a_step_kh + kw * a_step_kw + kb * a_step_kb;
              constexpr int nks_full = G::swb >> 5;
              const int nks = (kb == G::nkb - 1) ? G::ks_last : nks_full;
#pragma unroll
              for (int ks = 0; ks < nks; ++ks)
                if (!dbg_flag(p, 1)) umma_i8_pair(d0, a_desc + 2 * ks, b_desc + 2 * ks, idesc0, (blk | ks) != 0);
            }
            umma_commit_pair_local(smem_u32(&bar->g1_prog[tap_i % kG1Ahead]));
            ++tap_i;
          }
        }
        umma_commit_pair(smem_u32(&bar->a_empty[sa]));
        umma_commit_pair(smem_u32(&bar->acc0_full[ab]));
        tr.ev(11);
        if (++sa == (uint32_t)p.SA) { sa = 0; a_par ^= 1; }
      }
    }
   }
  } else if (warp >= kEpiWarp0) {
    // ================================== epilogue (both CTAs) ===================================
    load_epilogue_constants<G>(p, smem);
    if constexpr (static_two_groups<G>())
      epilogue_static2<G, kDst, true>(p, smem, bar, tmem, warp, lane, n_local, 2 * cid + (int)rank, 2 * ncl);
    else if constexpr (static_epilogue<G>())
      epilogue_static<G, kDst, true>(p, tmD, smem, bar, tmem, warp, lane, n_local, 2 * cid + (int)rank, 2 * ncl);
    else
      epilogue_role<G, kDst, false, false, false, true>(p, smem, bar, tmem, warp, lane, n_local, 2 * cid + (int)rank, 2 * ncl);
  } else if (warp == 3) {
    if constexpr (!G::is_static) {
      if ((kDst == DF_U8 || kDst == DF_S8) && p.stage_out && elect_one())
        store_role<G>(p, tmD, bar, sbase, n_local, 2 * cid + (int)rank, 2 * ncl);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still touch it
  if (warp == 3) tmem_dealloc_pair<512>(tmem);
  trace_wallclock(p, 1);
}

// ================================================================ launchers (host side, per instantiation)
// The kernel instantiations are spread over several translation units (conv_inst_*.cu) so that they compile
// in parallel; conv_fused.cu (create / run) reaches them through these type-erased function pointers.
// type-erased launcher: the epilogue-constant parameter type depends on the geometry
typedef cudaError_t (*LaunchFn)(int grid, uint32_t smem, cudaStream_t st, const SrcMaps& a, const CUtensorMap& w0,
                                const CUtensorMap& w1, const DstMaps& d, const Params& p);
typedef cudaError_t (*AttrFn)(uint32_t smem);

// Launch with programmatic stream serialization: back-to-back launches on one stream overlap the next
// launch's prologue with this launch's tail (see griddep_launch_dependents in the kernels).
// DF_NO_PDL=1 (read once per process) launches without the programmatic-stream-serialization attribute
bool pdl_enabled();

template <class Kernel, class AMaps>
cudaError_t launch_pdl(Kernel kernel, int grid, int threads, uint32_t smem, cudaStream_t st, const AMaps& a,
                       const CUtensorMap& w0, const CUtensorMap& w1, const DstMaps& d, const Params& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, a, w0, w1, d, p);
}

template <class G, int kDst, bool kDown0, bool kDown1, bool kNanSafe>
cudaError_t launch_conv(int grid, uint32_t smem, cudaStream_t st, const SrcMaps& a, const CUtensorMap& w0,
                        const CUtensorMap& w1, const DstMaps& d, const Params& p) {
  return launch_pdl(conv_fused_kernel<G, kDst, kDown0, kDown1, kNanSafe>, grid, kThreads, smem, st, a, w0, w1, d, p);
}
template <class G, int kDst, bool kDown0, bool kDown1, bool kNanSafe>
cudaError_t attr_conv(uint32_t smem) {
  return cudaFuncSetAttribute((const void*)conv_fused_kernel<G, kDst, kDown0, kDown1, kNanSafe>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
struct KernelFn {
  LaunchFn launch;
  AttrFn attr;
};

template <class G, int kDst>
cudaError_t launch_pair(int grid, uint32_t smem, cudaStream_t st, const SrcMaps& a, const CUtensorMap& w0,
                        const CUtensorMap& w1, const DstMaps& d, const Params& p) {
  return launch_pdl(conv_pair_kernel<G, kDst>, grid, kThreads, smem, st, a.m[0], w0, w1, d, p);
}
template <class G, int kDst>
cudaError_t attr_pair(uint32_t smem) {
  return cudaFuncSetAttribute((const void*)conv_pair_kernel<G, kDst>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
#define DF_KERNEL(G, DT, D0, D1, NS) KernelFn{launch_conv<G, DT, D0, D1, NS>, attr_conv<G, DT, D0, D1, NS>}

// Static geometries = the BASELINE.json conv shapes together with the shared-memory plan
// df_conv_create derives for them (weights resident? how many weight stages).
using GeoCfg1 = StaticGeom<64, 64, 256, 1, 1, 1>;      // 56x56  64->64->256 : everything resident
using GeoCfg3 = StaticGeom<128, 128, 512, 0, 1, 3>;    // 28x28 128->128->512: W1 resident, W0 through 3 stages
using GeoCfg4 = StaticGeom<256, 256, 1024, 0, 0, 2>;   // 14x14 256->256->1024: all weights through 2 stages
using GeoCfg3P = StaticGeom<128, 128, 512, 1, 1, 1>;   // cfg3 on CTA pairs: weight halves resident (conv_pair_kernel)
using GeoCfg4P = StaticGeom<256, 256, 1024, 0, 0, 2>;  // cfg4 on CTA pairs: weight halves streamed (conv_pair_kernel)

// defined in conv_inst_*.cu
KernelFn pick_pair_cfg3(int dst_dt);                                             // conv_pair_kernel<GeoCfg3P, dst>
KernelFn pick_pair_cfg4(int dst_dt);                                             // conv_pair_kernel<GeoCfg4P, dst>
KernelFn pick_pair_dyn(int dst_dt);                                              // conv_pair_kernel<DynGeom, dst>
KernelFn pick_static_geom(int geom_id, int dst_dt);                              // conv_fused_kernel<GeoCfg{1,3,4}, dst>
KernelFn pick_dynamic_u8(bool down0, bool down1, bool nan_safe);                 // conv_fused_kernel<DynGeom, ...>
KernelFn pick_dynamic_s8(bool down0, bool down1, bool nan_safe);
KernelFn pick_dynamic_s32(bool down0, bool down1, bool nan_safe);
KernelFn pick_dynamic_f32(bool down0, bool down1, bool nan_safe);

}  // namespace dfconv
