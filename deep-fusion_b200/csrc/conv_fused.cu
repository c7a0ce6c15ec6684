// conv_fused.cu -- host side of the fused conv: df_conv_create / df_conv_run / df_conv_query (C-ABI of
// include/dfcuda.h).  The kernels are in conv_kernels.cuh, instantiated in conv_inst_*.cu.
#include <vector>

#include "conv_kernels.cuh"

using namespace dfconv;

namespace dfconv {
KernelFn pick_static_cfg1(int dst_dt);
KernelFn pick_static_cfg3(int dst_dt);
KernelFn pick_static_cfg4(int dst_dt);
KernelFn pick_static_geom(int geom_id, int dst_dt) {
  if (geom_id == 1) return pick_static_cfg1(dst_dt);
  if (geom_id == 3) return pick_static_cfg3(dst_dt);
  if (geom_id == 4) return pick_static_cfg4(dst_dt);
  return KernelFn{nullptr, nullptr};
}
bool pdl_enabled() {
  static const bool on = getenv("DF_NO_PDL") == nullptr;
  return on;
}
}  // namespace dfconv

namespace {

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

// geom_id: 0 = run-time geometry, 1 / 3 / 4 = GeoCfg1 / GeoCfg3 / GeoCfg4
KernelFn pick_kernel(int geom_id, int dst_dt, bool down0, bool down1, bool nan_safe) {
  if (geom_id && !down0 && !down1 && !nan_safe) return pick_static_geom(geom_id, dst_dt);
  switch (dst_dt) {
    case DF_U8: return pick_dynamic_u8(down0, down1, nan_safe);
    case DF_S8: return pick_dynamic_s8(down0, down1, nan_safe);
    case DF_S32: return pick_dynamic_s32(down0, down1, nan_safe);
    case DF_F32: return pick_dynamic_f32(down0, down1, nan_safe);
  }
  return KernelFn{nullptr, nullptr};
}

template <class G>
bool geom_matches(const Params& p) {
  return p.IC == G::IC && p.OC == G::OC && p.OC1 == G::OC1 && p.w0_res == G::w0_res && p.w1_res == G::w1_res &&
         p.SB >= G::SB && p.swb == G::swb && p.nkb == G::nkb && p.swb1 == G::swb1 && p.nkb1 == G::nkb1 &&
         p.nc1 == G::nc1 && p.n_chunks == G::n_chunks && p.n_acc0 == G::n_acc0;
}

}  // namespace

struct df_conv {
  df_conv() : desc(), prm(), kernel{nullptr, nullptr}, pair(false), pair_kernel{nullptr, nullptr}, pair_prm(), pair_smem(0),
              tmW0h(), tmW1h(), geom_id(0), smem_bytes(0), device(0), sms(0), d_w0(nullptr),
              d_w1(nullptr), d_bias0(nullptr), d_scale0(nullptr), d_bias1(nullptr), d_scale1(nullptr), d_k1(nullptr),
              tmW0(), tmW1(), a_maps(), d_maps(), a_next(0), d_next(0), trace(nullptr), trace_cap(0), n_src(1), src_ic(), parts() {}
  df_conv_desc desc;
  Params prm;        // everything except n-dependent fields and dst
  KernelFn kernel;
  // CTA-pair variant (conv_pair_kernel), used when `pair` is set
  bool pair;
  KernelFn pair_kernel;
  Params pair_prm;
  uint32_t pair_smem;
  CUtensorMap tmW0h, tmW1h;
  int geom_id;
  uint32_t smem_bytes;
  int device, sms;
  int8_t *d_w0, *d_w1;
  float *d_bias0, *d_scale0, *d_bias1, *d_scale1;
  int* d_k1;
  CUtensorMap tmW0, tmW1;
  // activation / destination tensor maps depend on (pointer, batch): a small round-robin cache keeps
  // callers that cycle through a few buffers from re-encoding on every call
  struct SrcSlot {
    const void* ptr[kMaxSrc];  // ptr[0] == nullptr: empty slot
    int n;
    SrcMaps maps;
  };
  struct DstSlot {
    const void* ptr;
    int n;
    DstMaps maps;
  };
  static constexpr int kMapSlots = 16;
  SrcSlot a_maps[kMapSlots];
  DstSlot d_maps[kMapSlots];
  int a_next, d_next;
  unsigned long long* trace;
  int trace_cap;
  int n_src;            // inputs whose channel concatenation is the conv's source (1: plain conv)
  int src_ic[kMaxSrc];  // their channel counts
  // Composite operators (this handle then owns no kernel of its own):
  //   groups  : the conv-only operator with more than 256 output channels = one launch per group of <= 256
  //             channels, each writing its channel range of the same destination pixels;
  //   chained : the fused operator whose first stage has more than 256 output channels (or does not fit one
  //             CTA's shared memory) = conv-only stage -> u8 intermediate in device memory `d_mid` (L2-resident
  //             for the batches this is used with) -> 1x1 conv-only stage.  Same arithmetic, bit for bit: the
  //             fused kernel's intermediate IS the conv-only operator's u8 output.
  std::vector<df_conv*> parts;
  bool chained = false;
  void* d_mid = nullptr;
  bool with_sum = false;  // created by df_conv_create_sum: run-time-geometry kernel, direct stores
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

// accumulator column -> output channel for an accumulator of `ncols` real columns (a multiple of 16):
// blocks of 32 columns, then one block of 16; inside a block of NB columns, column 8k + 2m + e holds
// channel (NB/4) * m + 2k + e (see tmem_ld_16x256b_x8)
int col_to_channel(int col, int ncols) {
  const int full = ncols / 32 * 32;
  int base, nb;
  if (col < full) {
    base = col / 32 * 32;
    nb = 32;
  } else {
    base = full + (col - full) / 16 * 16;
    nb = 16;
  }
  const int r = col - base, k = r / 8, m = (r % 8) / 2, e = r % 2;
  return base + (nb / 4) * m + 2 * k + e;
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

static int conv_create_impl(const df_conv_desc* d, int n_src, const int* src_ic, int concat_relu, const int8_t* wei,
                            const int8_t* wei1, const void* bia0, const void* bia1, const float* scale0,
                            const float* scale1, df_conv** out, bool with_sum = false);

static size_t dt_bytes(int dt) { return (dt == DF_F32 || dt == DF_S32) ? 4 : 1; }

// Composite operators (see df_conv::parts).  Everything the reference accepts for these shapes
// (src/jit_conv_kernel.cc:586-661 has no upper bound on oc) runs on the device; nothing falls back to the CPU.
static int conv_create_composite(const df_conv_desc* d, const int8_t* wei, const int8_t* wei1, const void* bia0,
                                 const void* bia1, const float* scale0, const float* scale1, df_conv** out, bool with_sum) {
  df_conv* op = new df_conv();
  op->desc = *d;
  op->with_sum = with_sum;
  int rc = 0;
  if (d->oc1 == 0) {
    // ---- conv-only, oc > 256: groups of <= 256 output channels.  OIhw4i16o4i is output-block major, so the
    //      weights of channels [o0, o0 + n) are one contiguous slice (jit_conv_kernel.cc:333-338).
    for (int o0 = 0; o0 < d->oc && rc == 0; o0 += 256) {
      df_conv_desc g = *d;
      g.oc = d->oc - o0 < 256 ? d->oc - o0 : 256;
      g.nscale0 = d->nscale0 > 1 ? g.oc : 1;
      const int8_t* w = wei + (size_t)(o0 / 16) * (d->ic / 16) * d->kh * d->kw * 256;
      const void* b = bia0 ? static_cast<const char*>(bia0) + (size_t)o0 * dt_bytes(d->bia0_dt) : nullptr;
      df_conv* part = nullptr;
      rc = conv_create_impl(&g, 0, nullptr, 0, w, nullptr, b, nullptr, scale0 + (d->nscale0 > 1 ? o0 : 0), nullptr, &part, with_sum);
      if (rc) break;
      part->prm.dst_pitch = part->pair_prm.dst_pitch = d->oc;
      part->prm.dst_ch0 = part->pair_prm.dst_ch0 = o0;
      op->parts.push_back(part);
    }
  } else {
    // ---- fused, first stage too wide for one accumulator / one CTA's shared memory: conv-only stage with u8
    //      destination (ReLU + round0 + unsigned saturation = the fused kernel's intermediate, jit_conv_kernel.cc
    //      :264-277), then the 1x1 stage as a conv-only operator over that intermediate
    op->chained = true;
    df_conv_desc a = *d;
    a.oc1 = 0;
    a.dst_dt = DF_U8;
    a.bia1_dt = DF_UNDEF;
    a.relu0 = 1;
    df_conv_desc b = *d;
    b.ih = (d->ih + 2 * d->ph - d->kh) / d->sh + 1;
    b.iw = (d->iw + 2 * d->pw - d->kw) / d->sw + 1;
    b.ic = d->oc;
    b.oc = d->oc1;
    b.oc1 = 0;
    b.kh = b.kw = b.sh = b.sw = 1;
    b.ph = b.pw = 0;
    b.bia0_dt = d->bia1_dt;
    b.bia1_dt = DF_UNDEF;
    b.relu0 = d->relu1;
    b.round0 = d->round1;
    b.nscale0 = d->nscale1;
    df_conv *pa = nullptr, *pb = nullptr;
    if (!wei1 || !scale1) rc = df::fail(DF_E_INVALID, "conv: null 1x1 weights / scales");
    if (!rc) rc = conv_create_impl(&a, 0, nullptr, 0, wei, nullptr, bia0, nullptr, scale0, nullptr, &pa);
    if (pa) op->parts.push_back(pa);
    if (!rc) rc = conv_create_impl(&b, 0, nullptr, 0, wei1, nullptr, bia1, nullptr, scale1, nullptr, &pb, with_sum);
    if (pb) op->parts.push_back(pb);
    if (!rc) {
      cudaError_t e = cudaMalloc(&op->d_mid, (size_t)d->n * b.ih * b.iw * d->oc);
      if (e != cudaSuccess) rc = df::fail((int)e, "cudaMalloc(intermediate) failed: %s", cudaGetErrorString(e));
    }
  }
  if (rc) {
    df_conv_destroy(op);
    return rc;
  }
  cudaGetDevice(&op->device);
  op->n_src = 1;
  *out = op;
  return 0;
}

// n_src == 0: plain conv.  n_src >= 1: the source is the channel concatenation of n_src tensors (fused concat)
static int conv_create_impl(const df_conv_desc* d, int n_src, const int* src_ic, int concat_relu, const int8_t* wei,
                            const int8_t* wei1, const void* bia0, const void* bia1, const float* scale0,
                            const float* scale1, df_conv** out, bool with_sum) {
  if (!out) return df::fail(DF_E_INVALID, "conv: null out");
  *out = nullptr;
  int rc = validate(d);
  if (rc) return rc;
  const bool fused_cat = n_src > 0;
  int swb_cat = 128;
  if (fused_cat) {
    // acceptance of the concat half = jit_concat_kernel::init_conf for 1-byte types (every input's channel
    // count a multiple of 16, src/jit_concat_kernel.cc:157-176); the fused A-operand load additionally needs
    // whole 32-byte K-steps per input
    if (n_src > kMaxSrc || !src_ic) return df::fail(DF_E_UNSUPPORTED, "concat+conv: at most %d inputs", kMaxSrc);
    long sum = 0;
    for (int i = 0; i < n_src; ++i) {
      if (src_ic[i] <= 0 || src_ic[i] % 16) return df::fail(DF_E_INVALID, "concat+conv: input %d has %d channels (not a positive multiple of 16)", i, src_ic[i]);
      if (src_ic[i] % 32) return df::fail(DF_E_UNSUPPORTED, "concat+conv: fused A-operand load needs channel counts that are multiples of 32 (input %d: %d)", i, src_ic[i]);
      while (src_ic[i] % swb_cat) swb_cat /= 2;
      sum += src_ic[i];
    }
    if (sum != d->ic) return df::fail(DF_E_INVALID, "concat+conv: inputs have %ld channels, the conv expects %d", sum, d->ic);
    if (d->ic / swb_cat > kMaxKBlocks) return df::fail(DF_E_UNSUPPORTED, "concat+conv: more than %d halo K-blocks", kMaxKBlocks);
  }
  if (!wei || !scale0) return df::fail(DF_E_INVALID, "conv: null weights / scales");
  if ((d->bia0_dt != DF_UNDEF && !bia0) || (d->oc1 != 0 && d->bia1_dt != DF_UNDEF && !bia1))
    return df::fail(DF_E_INVALID, "conv: bias dtype given but pointer is null");
  // ---- the B200 path: fused 3x3 s1 p1 + 1x1
  const bool conv0_only = d->oc1 == 0;  // conv() without the 1x1 stage (include/deepfusion.h:121-129)
  if (!conv0_only && (!wei1 || !scale1)) return df::fail(DF_E_INVALID, "conv: null 1x1 weights / scales");
  // any window with stride 1 whose output is not larger than its input (2 p <= k - 1: every "same" or "valid"
  // convolution); taps are constant offsets in the linearised padded pixel space (conv_kernels.cuh)
  if (fused_cat && (d->sh != 1 || d->sw != 1)) return df::fail(DF_E_UNSUPPORTED, "concat+conv: k3 s1 p1 only");
  if (2 * d->ph > d->kh - 1 || 2 * d->pw > d->kw - 1)
    return df::fail(DF_E_UNSUPPORTED, "B200 path supports padding <= (kernel - 1) / 2 (got k %dx%d p %dx%d)", d->kh, d->kw, d->ph, d->pw);
  if (d->kh * d->kw > 121) return df::fail(DF_E_UNSUPPORTED, "B200 path supports windows of at most 121 taps");
  if (fused_cat && (d->kh != 3 || d->kw != 3 || d->ph != 1 || d->pw != 1))
    return df::fail(DF_E_UNSUPPORTED, "concat+conv: k3 s1 p1 only");
  if (d->oc > 256) {  // more channels than one TMEM accumulator: composite operator
    if (fused_cat) return df::fail(DF_E_UNSUPPORTED, "concat+conv: conv0 oc <= 256 (got %d)", d->oc);
    return conv_create_composite(d, wei, wei1, bia0, bia1, scale0, scale1, out, with_sum);
  }
  const int zr_w = d->pw > d->kw - 1 - d->pw ? d->pw : d->kw - 1 - d->pw;  // zero columns between rows
  if (fused_cat && d->iw + zr_w > 256) return df::fail(DF_E_UNSUPPORTED, "concat+conv: padded width <= 256");

  df_conv* op = new df_conv();
  op->desc = *d;
  Params& p = op->prm;
  p.H = d->ih;
  p.W = d->iw;
  p.IC = d->ic;
  p.OC = d->oc;
  p.OC1 = conv0_only ? d->oc : d->oc1;  // destination channels
  p.conv0_only = conv0_only ? 1 : 0;
  p.w1_pack = 1;
  p.g_interleave = 0;
#if DF_DIAG
  p.dbg_no_mma = getenv("DF_DEBUG_NO_MMA") ? atoi(getenv("DF_DEBUG_NO_MMA")) : 0;  // bit 0: no MMA, 1: no TMA stores, 2: no staging writes
#else
  p.dbg_no_mma = 0;
#endif
  p.swb = pick_swb(d->ic);  // weight K-blocks
  p.nkb = (d->ic + p.swb - 1) / p.swb;
  p.swa = fused_cat ? swb_cat : p.swb;  // halo K-blocks: never straddle two inputs of a fused concat
  p.nka = (d->ic + p.swa - 1) / p.swa;
  p.n_src = fused_cat ? n_src : 1;
  p.concat_relu = fused_cat && concat_relu;
  op->n_src = p.n_src;
  op->with_sum = with_sum;
  op->src_ic[0] = d->ic;
  if (fused_cat)
    for (int i = 0, kb = 0; i < n_src; ++i) {
      op->src_ic[i] = src_ic[i];
      for (int c0 = 0; c0 < src_ic[i]; c0 += p.swa, ++kb) {
        p.kb_src[kb] = (unsigned char)i;
        p.kb_c0[kb] = (unsigned short)c0;
      }
    }
  p.ks_last = (d->ic - (p.nkb - 1) * p.swb + 31) / 32;
  p.swb1 = pick_swb(d->oc);
  p.nkb1 = (d->oc + p.swb1 - 1) / p.swb1;
  p.ks1_last = (d->oc - (p.nkb1 - 1) * p.swb1 + 31) / 32;
  p.KH = d->kh;
  p.KW = d->kw;
  p.PH = d->ph;
  p.PW = d->pw;
  p.OH = d->ih + 2 * d->ph - d->kh + 1;
  p.OW = d->iw + 2 * d->pw - d->kw + 1;
  p.SH = d->sh;
  p.SW = d->sw;
  p.OHS = (d->ih + 2 * d->ph - d->kh) / d->sh + 1;
  p.OWS = (d->iw + 2 * d->pw - d->kw) / d->sw + 1;
  p.ZR = d->ph > d->kh - 1 - d->ph ? d->ph : d->kh - 1 - d->ph;
  p.Hp = d->ih + p.ZR;
  const int wp_align = 128 / p.swa;  // every halo row must start 128 B aligned for TMA
  p.Wp = (d->iw + zr_w + wp_align - 1) / wp_align * wp_align;
  p.n_box = (p.Wp + 255) / 256;  // TMA boxes are at most 256 positions wide
  p.box_w = ((p.Wp + p.n_box - 1) / p.n_box + wp_align - 1) / wp_align * wp_align;
  p.Wp = p.n_box * p.box_w;
  p.q_first = (p.ZR + 1) * p.Wp;
  // rows touched by the 128 + (KH - 1) * Wp + KW - 1 consecutive positions a tile's taps read
  p.NR = (kTileM + (p.KH - 1) * p.Wp + p.KW - 3) / p.Wp + 2;
  p.r8_dw = 8 % p.Wp;
  p.r8_dn = (8 / p.Wp) / p.Hp;
  p.r8_dh = (8 / p.Wp) % p.Hp;
  p.dst_pitch = p.OC1;
  p.dst_ch0 = 0;
  p.nc1 = p.OC1 < 128 ? p.OC1 : 128;
  p.n_chunks = (p.OC1 + p.nc1 - 1) / p.nc1;
  // two conv0 accumulators when both fit: always for oc <= 128; the conv-only operator has no conv1 accumulators in the
  // other half of the TMEM, so there also for oc <= 256 (GEMM1 of the next tile then runs under this tile's epilogue)
  p.n_acc0 = (d->oc <= 128 || (conv0_only && d->oc <= 256 && !getenv("DF_NO_C0_DOUBLE"))) ? 2 : 1;
  p.relu1 = conv0_only ? d->relu0 : d->relu1;
  p.round0 = d->round0;
  p.round1 = conv0_only ? d->round0 : d->round1;

  // ---- parameters: weights re-laid out K-major per (tap, K-block); bias -> f32; scales expanded
  // Row r of a weight block (= MMA N index = accumulator column r) holds output channel
  // col_to_channel(r): the order the epilogue's row-pair fragments want (see tmem_ld_16x256b_x8).
  const int taps = p.KH * p.KW;
  std::vector<int8_t> w0((size_t)taps * p.nkb * p.OC * p.swb, 0);
  for (int tap = 0; tap < taps; ++tap)
    for (int r = 0; r < p.OC; ++r) {
      const int o = col_to_channel(r, p.OC);
      for (int i = 0; i < p.IC; ++i) {
        const int kb = i / p.swb;
        w0[(((size_t)tap * p.nkb + kb) * p.OC + r) * p.swb + (i - kb * p.swb)] =
            wei[blocked_off(o, i, tap / p.KW, tap % p.KW, p.IC, p.KH, p.KW)];
      }
    }
  std::vector<int8_t> w1((size_t)p.n_chunks * p.nkb1 * p.nc1 * p.swb1, 0);  // stays zero (and unused) for conv0-only
  for (int j = 0; j < (conv0_only ? 0 : p.n_chunks); ++j) {
    const int ncols = p.OC1 - j * p.nc1 < p.nc1 ? p.OC1 - j * p.nc1 : p.nc1;  // real columns of this chunk
    for (int r = 0; r < ncols; ++r) {
      const int q = j * p.nc1 + col_to_channel(r, ncols);
      for (int o = 0; o < p.OC; ++o) {
        const int kb = o / p.swb1;
        w1[(((size_t)j * p.nkb1 + kb) * p.nc1 + r) * p.swb1 + (o - kb * p.swb1)] = wei1[blocked_off(q, o, 0, 0, p.OC, 1, 1)];
      }
    }
  }
  std::vector<float> b0(p.OC), s0(p.OC), b1(p.OC1), s1(p.OC1);
  bool finite = true;
  for (int o = 0; o < p.OC; ++o) {
    b0[o] = d->bia0_dt != DF_UNDEF ? bias_to_f32(d->bia0_dt, bia0, o) : 0.f;  // x + (+0.0f) == x here
    s0[o] = scale0[d->nscale0 > 1 ? o : 0];                                    // broadcast (defect D4)
    finite = finite && isfinite(b0[o]) && isfinite(s0[o]);
  }
  for (int q = 0; q < p.OC1; ++q) {
    if (conv0_only) {  // the final stage of the conv0-only operator uses the conv0 bias / scales
      b1[q] = b0[q];
      s1[q] = s0[q];
      continue;
    }
    b1[q] = d->bia1_dt != DF_UNDEF ? bias_to_f32(d->bia1_dt, bia1, q) : 0.f;
    s1[q] = scale1[d->nscale1 > 1 ? q : 0];
    finite = finite && isfinite(b1[q]) && isfinite(s1[q]);
  }
  p.nan_safe = !finite;

  // conv1 offset-magic conversion (see scale4_fast): exact iff the accumulator range of every
  // channel is narrower than 2^23 and C[q] = bias + lo - 2^23 is exactly representable.
  std::vector<int> k1(p.OC1, 0);
  std::vector<float> c1(p.OC1, 0.f);
  bool fast1 = finite && !conv0_only;  // (the 3x3 accumulator's range is far wider than 2^23)
  if (!conv0_only) {
    std::vector<long long> lo(p.OC1), hi(p.OC1);
    long long lo_min = 0, hi_max = 0;
    for (int q = 0; q < p.OC1; ++q) {
      long long neg = 0, pos = 0;
      for (int o = 0; o < p.OC; ++o) {
        const int w = wei1[blocked_off(q, o, 0, 0, p.OC, 1, 1)];
        if (w < 0) neg += w; else pos += w;
      }
      lo[q] = 255 * neg;
      hi[q] = 255 * pos;
      lo_min = lo[q] < lo_min ? lo[q] : lo_min;
      hi_max = hi[q] > hi_max ? hi[q] : hi_max;
    }
    // one K for every channel when a single lower bound keeps all ranges inside [0, 2^23)
    bool uniform = hi_max - lo_min < (1ll << 23) && !getenv("DF_NO_UNIFORM_K");  // (env: test hook, per-channel K path)
    for (int pass = 0; pass < 2; ++pass) {
      bool ok = finite;
      for (int q = 0; q < p.OC1 && ok; ++q) {
        const long long base = uniform ? lo_min : lo[q];
        if (hi[q] - base >= (1ll << 23)) { ok = false; break; }
        const double c = (double)b1[q] + (double)base - 8388608.0;
        if ((double)(float)c != c) { ok = false; break; }
        k1[q] = (int)(0x4B000000ll - base);
        c1[q] = (float)c;
      }
      if (ok || !uniform) { fast1 = ok; break; }
      uniform = false;  // retry with per-channel offsets
    }
    p.k1_uniform = (fast1 && uniform) ? k1[0] : 0;
  } else {
    p.k1_uniform = 0;
  }
  if (getenv("DF_NO_FAST_CONV1")) fast1 = false;  // test hook: exercise the I2F path
  if (!fast1) p.k1_uniform = 0;
  p.fast1 = fast1;

  // ---- shared memory plan.  plan(nm, stage): nm intermediate tiles; stage = staged 1-byte output wanted
  const int oc1_pad = p.n_chunks * p.nc1;
  const uint32_t avail = kSmemLimit - 1024;  // base alignment slack
  // sa_min: halo stages when the weights stream (2; 1 as a last resort for shapes whose halo alone is > 100 KB --
  // loads and MMAs of successive tiles then alternate instead of overlapping)
  int ks_split = 1;  // halo K-slices per tile (Params::n_ks): > 1 only when an unsliced halo stage cannot fit
  auto plan = [&](int nm, bool stage, int sa_min = 2) -> bool {
    p.n_ks = ks_split;
    p.nkb_s = p.nka / ks_split;
    uint32_t off = 1024;  // barriers
    p.off_bias0 = off;
    off += align_up(p.OC * 4, 128);
    p.off_scale0 = off;
    off += align_up(p.OC * 4, 128);
    p.off_bias1 = off;
    off += align_up(oc1_pad * 4, 128);
    p.off_scale1 = off;
    off += align_up(oc1_pad * 4, 128);
    p.off_k1 = off;
    off += align_up(oc1_pad * 4, 128);
    off = align_up(off, 1024);
    p.mid_kb_stride = kTileM * p.swb1;
    p.mid_bytes = align_up(p.nkb1 * p.mid_kb_stride, 1024);
    p.off_mid = off;
    p.NM = nm;
    off += nm * p.mid_bytes;
    p.a_kb_stride = (uint32_t)p.NR * p.Wp * p.swa;
    p.a_stage_bytes = align_up((p.nka / ks_split) * p.a_kb_stride, 1024);
    p.w0_block_bytes = (uint32_t)p.OC * p.swb;
    p.w1_block_bytes = (uint32_t)p.nc1 * p.swb1;
    const uint32_t w0_bytes = align_up(taps * p.nkb * p.w0_block_bytes, 1024);
    const uint32_t w1_bytes = conv0_only ? 0 : align_up(p.n_chunks * p.nkb1 * p.w1_block_bytes, 1024);
    const uint32_t fixed = off;
    if (ks_split == 1 && fixed + 2 * p.a_stage_bytes + w0_bytes + w1_bytes <= avail) {
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
      p.SA = sa_min;
      if (ks_split == 1 && fixed + p.SA * p.a_stage_bytes + w1_bytes + 3 * stage_w0 <= avail) {
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
      if (p.off_b + 2 * p.b_stage_bytes > avail) return false;
      int sb = (int)((avail - p.off_b) / p.b_stage_bytes);
      p.SB = sb > kMaxBStages ? kMaxBStages : sb;
    }
    // Staged output (store_staged_chunk, run-time geometry only): two 16 KB buffers behind the weight
    // stages; weight stages beyond four are given up for it.
    p.stage_out = 0;
    p.off_stage = 0;
    const uint32_t need = kStageBufs * kStageBytes;
    if (stage)
      while (p.SB > 4 && p.off_b + p.SB * p.b_stage_bytes + need > avail) --p.SB;
    uint32_t end = p.off_b + p.SB * p.b_stage_bytes;
    if (stage && end + need <= avail) {
      p.stage_out = 1;
      p.off_stage = end;
      end += need;
    }
    op->smem_bytes = end + 1024;
    return true;
  };
  // The static kernels (BASELINE shapes) are built for round-to-nearest, finite constants and the
  // offset-magic conv1 conversion with ONE constant K; they store 1-byte output straight from registers
  // and -- when it fits -- keep TWO intermediate tiles, so that the conv0 epilogue of tile t+1 does not have
  // to wait for GEMM2 of tile t (profiles/r02_knockout.log: with one tile the two strictly alternate and
  // their hand-offs alone cost ~3000 cycles per tile).  Anything else runs the run-time-geometry kernel.
  const bool static_ok = !conv0_only && taps == 9 && p.KH == 3 && p.PH == 1 && p.PW == 1 && p.SH * p.SW == 1 && p.n_box == 1 && p.fast1 && p.k1_uniform != 0 && d->round0 == DF_ROUND_NEAREST && d->round1 == DF_ROUND_NEAREST &&
                         !p.nan_safe && !fused_cat && !with_sum && !getenv("DF_FORCE_DYNAMIC_GEOMETRY");  // (env: test hook for the generic path)
  auto match_static = [&]() {
    if (geom_matches<GeoCfg1>(p)) { p.SB = GeoCfg1::SB; return 1; }
    if (geom_matches<GeoCfg3>(p)) { p.SB = GeoCfg3::SB; return 3; }
    if (geom_matches<GeoCfg4>(p)) { p.SB = GeoCfg4::SB; return 4; }
    return 0;
  };
  op->geom_id = 0;
  if (static_ok) {
    for (int nm = 2; nm >= 1 && !op->geom_id; --nm)
      if (plan(nm, false)) {
        op->geom_id = match_static();
        if (op->geom_id) op->smem_bytes = p.off_b + p.SB * p.b_stage_bytes + 1024;
      }
  }
  const int shape_id = op->geom_id;  // which BASELINE shape this is (0: none)
  if (!op->geom_id) {
    const bool can_stage = (d->dst_dt == DF_U8 || d->dst_dt == DF_S8) && p.nc1 == 128 && !with_sum &&
                           !(getenv("DF_NO_STAGED_STORE") && atoi(getenv("DF_NO_STAGED_STORE")) != 0);
    bool planned = plan(1, can_stage) || plan(1, can_stage, 1) || plan(1, false, 1);
    // an input too deep for any halo stage: K-slice it (the fewest slices that fit two stages); not for a fused concat
    for (int ks = 2; !planned && !fused_cat && ks <= p.nka; ++ks)
      if (p.nka % ks == 0) {
        ks_split = ks;
        planned = plan(1, can_stage) || plan(1, false);
      }
    if (!planned) {
      ks_split = 1;
      delete op;
      if (!conv0_only && !fused_cat)  // the two stages as two launches need less shared memory each
        return conv_create_composite(d, wei, wei1, bia0, bia1, scale0, scale1, out, with_sum);
      return df::fail(DF_E_UNSUPPORTED, "conv: shape does not fit the shared-memory plan");
    }
  }

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
  DF_TRY_CUDA(cudaMalloc(&op->d_k1, p.OC1 * 4));

  DF_TRY_CUDA(cudaMemcpy(op->d_k1, k1.data(), p.OC1 * 4, cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_bias1, fast1 ? c1.data() : b1.data(), p.OC1 * 4, cudaMemcpyHostToDevice));
  DF_TRY_CUDA(cudaMemcpy(op->d_scale1, s1.data(), p.OC1 * 4, cudaMemcpyHostToDevice));
  p.bias0 = op->d_bias0;
  p.scale0 = op->d_scale0;
  p.bias1 = op->d_bias1;
  p.scale1 = op->d_scale1;
  p.k1 = op->d_k1;
  DF_TRY(encode_2d(&op->tmW0, op->d_w0, p.swb, (long)taps * p.nkb * p.OC, p.OC));
  DF_TRY(encode_2d(&op->tmW1, op->d_w1, p.swb1, (long)p.n_chunks * p.nkb1 * p.nc1, p.nc1));

  op->kernel = pick_kernel(op->geom_id, d->dst_dt, p.round0 == DF_ROUND_DOWN, p.round1 == DF_ROUND_DOWN, p.nan_safe != 0);
  if (!op->kernel.launch) {
    df_conv_destroy(op);
    return df::fail(DF_E_UNSUPPORTED, "conv: no kernel for this configuration in this build");
  }
  // the opt-in limit is a property of the kernel FUNCTION, shared by every handle that uses it: always raise it to
  // the maximum (a later handle with a smaller plan must not lower it under an earlier handle's launches)
  DF_TRY_CUDA(op->kernel.attr(kSmemLimit));

  // ---- CTA-pair variants (conv_pair_kernel; DF_PAIR=0 selects the single-CTA kernels).
  //   cfg3: everything resident once the weights are split in two -- the single-CTA kernel has to stream 144 KB of
  //         3x3 weights per 128-position tile, more than the L2 delivers at tensor-pipe speed;
  //   cfg4: 832 KB of weights cannot be resident, but streaming HALVES per CTA halves the L2 traffic per tile and
  //         doubles what one ring stage feeds (DF_PAIR4=0 keeps the single-CTA kernel for this shape only).
  //   run-time geometry: any shape whose weights have to stream (w0_res == 0) runs the streamed pair form too, for the
  //         same reason as cfg4 (DF_PAIR_DYN=0 keeps the single-CTA kernel); round-to-nearest and finite constants
  //         only (the pair kernel instantiates one rounding variant), one TMA box per halo row, 128-byte K-blocks
  //         (ic > 64), no fused concat.
  const bool pair_on = !(getenv("DF_PAIR") && atoi(getenv("DF_PAIR")) == 0);
  const bool pair4_on = pair_on && !(getenv("DF_PAIR4") && atoi(getenv("DF_PAIR4")) == 0);
  const bool pair_dyn = pair_on && !(getenv("DF_PAIR_DYN") && atoi(getenv("DF_PAIR_DYN")) == 0) && op->geom_id == 0 && !fused_cat &&
                        p.w0_res == 0 && p.n_box == 1 && p.n_ks == 1 && !p.nan_safe && p.round0 == DF_ROUND_NEAREST && p.round1 == DF_ROUND_NEAREST &&
                        p.swb == 128;  // (the pair kernel shifts the TMA destination by whole positions: 128-byte K-blocks only)
  if ((((shape_id == 3 && pair_on) || (shape_id == 4 && pair4_on)) && static_ok) || pair_dyn) {
    Params q = p;
    const bool resident = shape_id == 3 && !pair_dyn;
    q.w0_res = q.w1_res = resident ? 1 : 0;
    q.stage_out = 0;
    q.off_stage = 0;
    q.NM = resident ? 2 : 1;
    uint32_t off2 = p.off_mid + q.NM * q.mid_bytes;
    if (resident) {
      q.off_w0 = off2;
      off2 += align_up(9 * q.nkb * (q.w0_block_bytes / 2), 1024);
      q.off_w1 = off2;
      off2 += align_up(q.n_chunks * q.nkb1 * (q.w1_block_bytes / 2), 1024);
    }
    q.off_a = off2;
    q.a_kb_stride = (uint32_t)(q.NR + 1) * q.Wp * q.swb;  // one extra row of slack before the tile origin
    q.a_stage_bytes = align_up(q.nkb * q.a_kb_stride, 1024);
    bool ok;
    if (resident) {
      const int sa = (int)((avail - q.off_a) / q.a_stage_bytes);
      ok = sa >= 2;
      q.SA = sa > kMaxAStages ? kMaxAStages : sa;
      q.SB = 1;
      q.off_b = q.off_a + q.SA * q.a_stage_bytes;
      q.b_stage_bytes = 0;
      op->pair_smem = q.off_b + 1024;
    } else {
      q.SA = getenv("DF_PAIR_SA") ? atoi(getenv("DF_PAIR_SA")) : 2;  // (development knob)
      q.off_b = q.off_a + q.SA * q.a_stage_bytes;
      const uint32_t half0 = q.w0_block_bytes / 2, half1 = q.w1_block_bytes / 2;
      q.b_stage_bytes = align_up(half0 > half1 ? half0 : half1, 1024);
      q.g_interleave = (!conv0_only && !getenv("DF_NO_INTERLEAVE")) ? 1 : 0;
      q.w1_pack = (!conv0_only && half1 % 1024 == 0 && 2 * half1 <= q.b_stage_bytes && q.nkb1 % 2 == 0 && !getenv("DF_NO_W1_PACK")) ? 2 : 1;
      const int sb = q.off_b < avail ? (int)((avail - q.off_b) / q.b_stage_bytes) : 0;
      ok = sb >= 2;
      q.SB = sb > kMaxBStages ? kMaxBStages : sb;
      uint32_t end = q.off_b + q.SB * q.b_stage_bytes;
      if (pair_dyn) {  // staged 1-byte output (generic epilogue) when two staging buffers fit behind >= 3 ring stages
        const bool want_stage = (d->dst_dt == DF_U8 || d->dst_dt == DF_S8) && q.nc1 == 128 && !with_sum &&
                                !(getenv("DF_NO_STAGED_STORE") && atoi(getenv("DF_NO_STAGED_STORE")) != 0);
        const uint32_t need = kStageBufs * kStageBytes;
        if (want_stage) {
          while (q.SB > 3 && q.off_b + q.SB * q.b_stage_bytes + need > avail) --q.SB;
          end = q.off_b + q.SB * q.b_stage_bytes;
          if (end + need <= avail) {
            q.stage_out = 1;
            q.off_stage = end;
            end += need;
          }
        }
      }
      op->pair_smem = end + 1024;
    }
    if (ok) {
      op->pair_prm = q;
      op->pair_kernel = pair_dyn ? pick_pair_dyn(d->dst_dt) : (resident ? pick_pair_cfg3(d->dst_dt) : pick_pair_cfg4(d->dst_dt));
      DF_TRY_CUDA(op->pair_kernel.attr(kSmemLimit));
      DF_TRY(encode_2d(&op->tmW0h, op->d_w0, q.swb, (long)taps * q.nkb * q.OC, q.OC / 2));
      DF_TRY(encode_2d(&op->tmW1h, op->d_w1, q.swb1, (long)q.n_chunks * q.nkb1 * q.nc1, q.nc1 / 2));
      op->pair = true;
    }
  }
  *out = op;
  return 0;
}

extern "C" int df_conv_create(const df_conv_desc* d, const int8_t* wei, const int8_t* wei1, const void* bia0,
                              const void* bia1, const float* scale0, const float* scale1, df_conv** out) {
  return conv_create_impl(d, 0, nullptr, 0, wei, wei1, bia0, bia1, scale0, scale1, out);
}

extern "C" int df_conv_create_concat(const df_conv_desc* d, int n_src, const int* src_ic, int concat_relu,
                                     const int8_t* wei, const int8_t* wei1, const void* bia0, const void* bia1,
                                     const float* scale0, const float* scale1, df_conv** out) {
  if (n_src <= 0) return df::fail(DF_E_INVALID, "concat+conv: no inputs");
  return conv_create_impl(d, n_src, src_ic, concat_relu, wei, wei1, bia0, bia1, scale0, scale1, out);
}

static int tiles_for(const Params& p, int n) {
  const long q_first = p.q_first, q_end = ((long)n * p.Hp + 1) * p.Wp;
  return (int)((q_end - q_first + kTileM - 1) / kTileM);
}

// Activation maps: 4-D u8 {C_i, W, H, n} over each NHWC source (one for a plain conv, n_src for a fused
// concat), box {swb, Wp, 1, 1} (one halo row of one K-block).  Looked up in / added to a round-robin cache keyed
// by (pointers, batch).
static int src_maps(df_conv* op, const Params& p, const void* const* ptrs, int n, const SrcMaps** out) {
  for (int i = 0; i < df_conv::kMapSlots; ++i) {
    const df_conv::SrcSlot& c = op->a_maps[i];
    if (!c.ptr[0] || c.n != n) continue;
    bool same = true;
    for (int k = 0; k < op->n_src; ++k) same = same && c.ptr[k] == ptrs[k];
    if (same) {
      *out = &c.maps;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return df::fail(DF_E_NODRIVER, "cuTensorMapEncodeTiled unavailable");
  df_conv::SrcSlot& s = op->a_maps[op->a_next];
  op->a_next = (op->a_next + 1) % df_conv::kMapSlots;
  s.ptr[0] = nullptr;
  for (int k = 0; k < op->n_src; ++k) {
    const cuuint64_t c = (cuuint64_t)op->src_ic[k];
    cuuint64_t gd[4] = {c, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)n};
    cuuint64_t gs[3] = {c, (cuuint64_t)p.W * c, (cuuint64_t)p.H * p.W * c};
    cuuint32_t box[4] = {(cuuint32_t)p.swa, (cuuint32_t)p.box_w, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&s.maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(ptrs[k]), gd, gs, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(p.swa), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return df::fail(DF_E_INTERNAL, "cuTensorMapEncodeTiled(src %d) failed: %d", k, (int)r);
  }
  for (int k = op->n_src - 1; k >= 0; --k) s.ptr[k] = ptrs[k];
  s.n = n;
  *out = &s.maps;
  return 0;
}

// Destination maps of the staged output path: the 1-byte NHWC destination as 2-D [n*H*W pixels][OC1],
// boxes {128 channels, 128 >> i pixels}, SWIZZLE_128B (store_staged_chunk).
// (The static-geometry kernels store per warp, 16 rows at most: they use m[3 ..].)
static int dst_maps(df_conv* op, const Params& p, const void* ptr, int n, const DstMaps** out) {
  const bool per_warp = false;
  for (int i = 0; i < df_conv::kMapSlots; ++i)
    if (op->d_maps[i].ptr == ptr && op->d_maps[i].n == n) {
      *out = &op->d_maps[i].maps;
      return 0;
    }
  EncodeTiledFn enc = get_encode();
  if (!enc) return df::fail(DF_E_NODRIVER, "cuTensorMapEncodeTiled unavailable");
  df_conv::DstSlot& s = op->d_maps[op->d_next];
  op->d_next = (op->d_next + 1) % df_conv::kMapSlots;
  s.ptr = nullptr;
  const cuuint64_t pixels = (cuuint64_t)n * p.OHS * p.OWS;
  for (int i = 0; i < (per_warp ? 6 : 8); ++i) {
    cuuint64_t gd[2] = {(cuuint64_t)p.dst_pitch, pixels};
    cuuint64_t gs[1] = {(cuuint64_t)p.dst_pitch};
    cuuint32_t box[2] = {per_warp ? 64u : 128u, (cuuint32_t)((per_warp ? 32 : 128) >> i)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&s.maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gd, gs, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, per_warp ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return df::fail(DF_E_INTERNAL, "cuTensorMapEncodeTiled(dst) failed: %d", (int)r);
  }
  s.ptr = ptr;
  s.n = n;
  *out = &s.maps;
  return 0;
}

static int conv_run_impl(df_conv* op, const void* const* srcs, void* dst, int n, void* stream, const void* res = nullptr) {
  if (op && res && !op->with_sum) return df::fail(DF_E_INVALID, "conv run: handle was not created for an eltwise sum (df_conv_create_sum)");
  if (res && (reinterpret_cast<uintptr_t>(res) & 15)) return df::fail(DF_E_INVALID, "conv run: residual must be 16-byte aligned");
  if (!op || !srcs || !dst) return df::fail(DF_E_INVALID, "conv run: null argument");
  if (!op->parts.empty()) {  // composite operator: see df_conv::parts
    if (n < 0 || n > op->desc.n) return df::fail(DF_E_INVALID, "conv run: batch %d outside [0, %d]", n, op->desc.n);
    if (op->chained) {
      const void* mid[1] = {op->d_mid};
      int rc = conv_run_impl(op->parts[0], srcs, op->d_mid, n, stream);
      return rc ? rc : conv_run_impl(op->parts[1], mid, dst, n, stream, res);
    }
    for (df_conv* part : op->parts) {
      int rc = conv_run_impl(part, srcs, dst, n, stream, res);
      if (rc) return rc;
    }
    return 0;
  }
  for (int k = 0; k < op->n_src; ++k)
    if (!srcs[k] || (reinterpret_cast<uintptr_t>(srcs[k]) & 15)) return df::fail(DF_E_INVALID, "conv run: source %d is null or not 16-byte aligned", k);
  if (n < 0 || n > op->desc.n) return df::fail(DF_E_INVALID, "conv run: batch %d outside [0, %d]", n, op->desc.n);
  if (n == 0) return 0;
  if (reinterpret_cast<uintptr_t>(dst) & 15) return df::fail(DF_E_INVALID, "conv run: src/dst must be 16-byte aligned");
  {  // the handle's weights, constants and tensor maps live on the device it was created on
    int dev = -1;
    DF_CUDA(cudaGetDevice(&dev));
    if (dev != op->device)
      return df::fail(DF_E_INVALID, "conv run: handle was created on device %d, current device is %d", op->device, dev);
  }
  Params p = op->pair ? op->pair_prm : op->prm;
  if ((long)n * p.Hp * p.Wp + 4L * p.Wp + kTileM >= (1L << 31))
    return df::fail(DF_E_UNSUPPORTED, "conv run: batch too large for 32-bit position index");
  const SrcMaps* tmA = nullptr;
  const DstMaps* tmD = &op->d_maps[0].maps;  // not read by the kernel unless stage_out
  {
    int rc = src_maps(op, p, srcs, n, &tmA);
    if (rc) return rc;
    if (p.stage_out) rc = dst_maps(op, p, dst, n, &tmD);
    if (rc) return rc;
  }
  p.N = n;
  p.n_tiles = tiles_for(p, n);
  p.dst = dst;
  p.res = res;
  p.trace = op->trace;
  p.trace_cap = op->trace_cap;
  auto set_tile_step = [&p](long tiles) {  // positions between successive tiles of one CTA
    const long dq = tiles * kTileM, rows = dq / p.Wp;
    p.ts_dw = (int)(dq % p.Wp);
    p.ts_dn = (int)(rows / p.Hp);
    p.ts_dh = (int)(rows % p.Hp);
  };
  if (op->pair) {
    const int pair_tiles = (p.n_tiles + 1) / 2, clusters = pair_tiles < op->sms / 2 ? pair_tiles : op->sms / 2;
    p.tile_step_mod = 0;
    set_tile_step(2L * clusters);
    DF_CUDA(op->pair_kernel.launch(2 * clusters, op->pair_smem, (cudaStream_t)stream, *tmA, op->tmW0h, op->tmW1h, *tmD, p));
    return 0;
  }
  const int grid = p.n_tiles < op->sms ? p.n_tiles : op->sms;
  p.tile_step_mod = (kTileM * grid) % p.Wp;
  set_tile_step(grid);
  DF_CUDA(op->kernel.launch(grid, op->smem_bytes, (cudaStream_t)stream, *tmA, op->tmW0, op->tmW1, *tmD, p));
  return 0;
}

extern "C" int df_conv_run(df_conv* op, const uint8_t* src, void* dst, int n, void* stream) {
  if (op && op->n_src != 1) return df::fail(DF_E_INVALID, "conv run: handle was created for %d concatenated inputs (df_conv_run_concat)", op->n_src);
  const void* srcs[1] = {src};
  return conv_run_impl(op, srcs, dst, n, stream);
}

extern "C" int df_conv_create_sum(const df_conv_desc* d, const int8_t* wei, const int8_t* wei1, const void* bia0,
                                  const void* bia1, const float* scale0, const float* scale1, df_conv** out) {
  return conv_create_impl(d, 0, nullptr, 0, wei, wei1, bia0, bia1, scale0, scale1, out, true);
}

extern "C" int df_conv_run_sum(df_conv* op, const uint8_t* src, const void* residual, void* dst, int n, void* stream) {
  if (!residual) return df::fail(DF_E_INVALID, "conv run: null residual");
  if (op && op->n_src != 1) return df::fail(DF_E_INVALID, "conv run: concat handle");
  const void* srcs[1] = {src};
  return conv_run_impl(op, srcs, dst, n, stream, residual);
}

extern "C" int df_conv_run_concat(df_conv* op, const void* const* srcs, void* dst, int n, void* stream) {
  return conv_run_impl(op, srcs, dst, n, stream);
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
  if (!op->parts.empty()) {  // composite: the first launch's plan, the whole operator's work
    int rc = df_conv_query(op->parts[0], info);
    double macs = 0;
    for (const df_conv* part : op->parts) {
      df_conv_info pi = {};
      if (!rc) rc = df_conv_query(part, &pi);
      macs += pi.macs_per_image;
    }
    info->macs_per_image = macs;
    return rc;
  }
  const Params& p = op->prm;
  info->tiles_per_launch = tiles_for(p, op->desc.n);
  info->grid = info->tiles_per_launch < op->sms ? info->tiles_per_launch : op->sms;
  info->block = kThreads;
  info->smem_bytes = (int)(op->pair ? op->pair_smem : op->smem_bytes);
  info->w0_resident = op->pair ? (op->pair_prm.w0_res ? 2 : 3) : p.w0_res;  // 2 = resident, split across a CTA pair; 3 = halves streamed by a CTA pair
  info->w1_resident = op->pair ? (op->pair_prm.w1_res ? 2 : 3) : p.w1_res;
  info->a_stages = op->pair ? op->pair_prm.SA : p.SA;
  info->b_stages = op->pair ? (op->pair_prm.w0_res ? 0 : op->pair_prm.SB) : p.SB;
  info->padded_w = p.Wp;
  info->padded_h = p.Hp;
  info->macs_per_image = (double)p.OHS * p.OWS * ((double)p.KH * p.KW * p.IC * p.OC + (p.conv0_only ? 0.0 : (double)p.OC * p.OC1));
  info->mma_efficiency = (double)op->desc.n * p.OHS * p.OWS / ((double)info->tiles_per_launch * kTileM);
  return 0;
}

extern "C" int df_conv_destroy(df_conv* op) {
  if (!op) return 0;
  for (df_conv* part : op->parts) df_conv_destroy(part);
  cudaFree(op->d_mid);
  cudaFree(op->d_w0);
  cudaFree(op->d_w1);
  cudaFree(op->d_bias0);
  cudaFree(op->d_scale0);
  cudaFree(op->d_bias1);
  cudaFree(op->d_scale1);
  cudaFree(op->d_k1);
  delete op;
  return 0;
}
