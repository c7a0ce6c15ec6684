/* ref_driver.cc -- drives the REFERENCE's own, unmodified kernel generators.  TEST INFRASTRUCTURE
 * (oracle/_ref/libdfref.so); never linked into or called from the product.
 *
 * oracle/Makefile compiles this file together with the reference sources where they lie under
 * /root/reference (src/jit_conv_kernel.cc, src/jit_concat_kernel.cc, src/op_concat.cc, src/op_conv.cc,
 * src/deepfusion.cc, util/x.cc) against oracle/xbyak_shim/xbyak/xbyak.h, a recording stand-in for the
 * un-vendored Xbyak: the reference's generate() runs as written, the instruction stream it emits is
 * interpreted with the host's own AVX-512 instructions.  What comes out is what the reference computes.
 *
 *   dfref_concat : the complete reference path, nothing of ours in between --
 *                  deepfusion::concat() -> op_concat<T> -> jit_concat_kernel (src/deepfusion.cc:105-121,
 *                  src/op_concat.cc:22-72, src/jit_concat_kernel.cc:30-197), driven through op::submit().
 *   dfref_conv   : jit_conv_kernel::init_conf + generate() unmodified (src/jit_conv_kernel.cc:27-673); the
 *                  HOST loop nest around the kernel is restated here, because the reference's own
 *                  (src/op_conv.cc:31-260) cannot run any real shape:
 *       D1  init_conf takes jcp.oc from the DESTINATION's channels and compares jcp.oh / jcp.ow with the
 *           1x1 weight's spatial dims (jit_conv_kernel.cc:577, :609-613), while op_conv::init_conf demands
 *           those be 1 (op_conv.cc:334): only oh = ow = 1, oc1x1 = oc passes both.  Here the kernel-level
 *           init_conf is called directly with DESCRIPTOR-ONLY memory objects shaped so that it derives the
 *           intended configuration (dst descriptor with conv0's oc channels, 1x1-weight descriptor with
 *           spatial dims oh x ow); every blocking decision (nb_ic_blocking, nb_oc_blocking, ur_w, use_vnni)
 *           is the reference's.
 *       D2  the driver's pointer arithmetic lacks factors (src row offset without *ic, weight strides
 *           without *256, bias offset times oc: op_conv.cc:165-166, :196-206).  The offsets below are the
 *           ones the emitted code's own addressing implies (input_offset / kernel_offset lambdas,
 *           jit_conv_kernel.cc:327-338; workspace layouts :27-48, :133-139, :193-216).
 *       D4  a single scale is read as a 64-byte vector (vmulps zword, :100, :263): the driver passes a
 *           16-wide replicated buffer.
 *       D7  scales are kept alive for the duration of the call (op_conv.h:93-95 stores a dangling pointer).
 *     D3 (fused + f32 destination saturates float bit patterns, :267, :275-277) is IN the generator and is
 *     therefore reproduced here; the oracle matches it with literal_f32_intermediate = 1.
 */
#include <algorithm>
#include <vector>

#include "deepfusion.h"
#include "deepfusion_utils.h"
#include "jit_concat_kernel.h"
#include "jit_conv_kernel.h"

#include "df_oracle.h"

using namespace deepfusion;

namespace {

typedef std::unique_ptr<memory> mem_p;

memory::dtype to_dtype(int dt) {
  switch (dt) {
    case DFO_F32: return memory::dtype::f32;
    case DFO_S32: return memory::dtype::s32;
    case DFO_S8: return memory::dtype::s8;
    case DFO_U8: return memory::dtype::u8;
    default: return memory::dtype::undef;
  }
}

struct ScaleBuf {  // D4: at least 16 floats behind the pointer the kernel reads a zword from
  std::vector<float> v;
  ScaleBuf(const float* s, int n) {
    if (n == 1) v.assign(16, s[0]);
    else v.assign(s, s + n);
  }
};

}  // namespace

extern "C" int dfref_supported(void) {
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
         __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vnni");
}

/* number of instructions the reference emitted for the last conv kernel built (diagnostic) */
static size_t g_last_conv_instructions = 0;
extern "C" long dfref_last_kernel_instructions(void) { return (long)g_last_conv_instructions; }

/* blocking the reference's init_conf picked for a shape: out = {nb_ic_blocking, nb_oc_blocking, ur_w, ur_w_tail, use_vnni} */
static int g_last_blocking[5];
extern "C" void dfref_last_blocking(int out[5]) { memcpy(out, g_last_blocking, sizeof(g_last_blocking)); }

extern "C" int dfref_concat(int dt, int relu, int n_inputs, const void* const* srcs, const int* ic, void* dst, long n_pixels) {
  if (!dfref_supported()) return -100;
  const memory::dtype mdt = to_dtype(dt);
  if (mdt == memory::dtype::undef || n_inputs <= 0 || n_pixels <= 0) return -1;
  const size_t ts = utils::dtype_size(mdt);
  std::vector<mem_p> ms;
  int oc = 0;
  for (int i = 0; i < n_inputs; ++i) {
    // logical (n, c, h, w) = (1, ic, 1, n_pixels): the op is per pixel (op_concat.cc:28, :58-60)
    ms.emplace_back(new memory(memory::nchw_dims{{1, ic[i], 1, (int)n_pixels}}, memory::format::nhwc, mdt));
    memcpy(ms.back()->data(), srcs[i], (size_t)n_pixels * ic[i] * ts);
    oc += ic[i];
  }
  mem_p md(new memory(memory::nchw_dims{{1, oc, 1, (int)n_pixels}}, memory::format::nhwc, mdt));
  jit::jit_concat_conf_t conf;
  if (!jit::jit_concat_kernel::init_conf(conf, ms, md, relu != 0)) return -2;  // (the op itself would exit(1))
  std::unique_ptr<op> o = concat(ms, md, relu != 0);
  o->submit();
  memcpy(dst, md->data(), (size_t)n_pixels * oc * ts);
  return 0;
}

extern "C" int dfref_conv(const dfo_conv_desc* d, const uint8_t* src, const int8_t* wei, const void* bia0, const float* scale0,
                          const int8_t* wei1, const void* bia1, const float* scale1, void* dst) {
  if (!dfref_supported()) return -100;
  const bool fused = d->oc1 != 0;
  const int oh = utils::conv_output_size(d->ih, d->kh, d->sh, d->ph), ow = utils::conv_output_size(d->iw, d->kw, d->sw, d->pw);
  if (oh <= 0 || ow <= 0) return -1;
  const memory::dtype dst_dt = to_dtype(d->dst_dt);

  // ---- descriptors for the kernel-level init_conf (D1: shaped so that it derives the intended jcp)
  mem_p m_src(new memory(memory::nchw_dims{{d->n, d->ic, d->ih, d->iw}}, memory::format::nhwc, memory::dtype::u8));
  mem_p m_wei(new memory(memory::nchw_dims{{d->oc, d->ic, d->kh, d->kw}}, memory::format::OIhw4i16o4i, memory::dtype::s8));
  mem_p m_dst(new memory(memory::nchw_dims{{d->n, d->oc, oh, ow}}, memory::format::nhwc, dst_dt));  // channels = conv0's oc
  mem_p m_bia, m_wei1, m_bia1;
  if (d->bia0_dt != DFO_UNDEF) m_bia.reset(new memory(memory::dims{d->oc}, memory::format::x, to_dtype(d->bia0_dt)));
  if (fused) {
    m_wei1.reset(new memory(memory::nchw_dims{{d->oc1, d->oc, oh, ow}}, memory::format::OIhw4i16o4i, memory::dtype::s8));
    if (d->bia1_dt != DFO_UNDEF) m_bia1.reset(new memory(memory::dims{d->oc1}, memory::format::x, to_dtype(d->bia1_dt)));
  }
  std::vector<float> sc0(scale0, scale0 + d->nscale0), sc1;
  if (fused) sc1.assign(scale1, scale1 + d->nscale1);
  else sc1.assign(1, 1.f);
  jit::jit_conv_conf_t jcp;
  if (!jit::jit_conv_kernel::init_conf(jcp, m_src, m_wei, m_bia, 1, {{d->sh, d->sw}}, {{d->ph, d->pw}}, m_dst, sc0, sc1, m_wei1,
                                       m_bia1, d->relu0 != 0, d->relu1 != 0, d->round0 == DFO_DOWN ? round_mode::down : round_mode::nearest,
                                       d->round1 == DFO_DOWN ? round_mode::down : round_mode::nearest))
    return -2;
  if (jcp.oc != d->oc || jcp.oh != oh || jcp.ow != ow || (fused && jcp.oc1x1 != d->oc1)) return -3;
  m_src.reset();  // descriptors only: release the (untouched) buffers
  m_wei.reset();
  m_dst.reset();
  m_wei1.reset();

  jit::jit_conv_kernel kernel(jcp);  // generate(): the reference's code generator runs here, unmodified
  g_last_conv_instructions = kernel.instructionCount();
  const auto& j = kernel.jcp;
  g_last_blocking[0] = j.nb_ic_blocking;
  g_last_blocking[1] = j.nb_oc_blocking;
  g_last_blocking[2] = j.ur_w;
  g_last_blocking[3] = j.ur_w_tail;
  g_last_blocking[4] = j.use_vnni ? 1 : 0;

  const ScaleBuf s0(scale0, d->nscale0), s1(fused ? scale1 : s0.v.data(), fused ? d->nscale1 : 1);
  const char* bias = static_cast<const char*>(bia0);
  char* out = static_cast<char*>(dst);
  const int oc_chunks = j.nb_oc / j.nb_oc_blocking, ic_chunks = j.nb_ic / j.nb_ic_blocking;
  const size_t blk = 256;  // one 16o x 16i block of OIhw4i16o4i
  const size_t ws_row = (size_t)j.ow * j.oc_block * j.nb_oc_blocking;  // s32 partials of one output row (conv0)
  const size_t ws1_row = (size_t)j.ow * (fused ? j.oc1x1 : 0);          // (oc1x1/16, ow, 16) per row (conv1)

  if (fused) {
    // corrected restatement of op_conv<T>::infer_conv0conv1 (src/op_conv.cc:140-260): rows of (n, oh) split
    // over threads, loops oc-chunk -> ic-chunk -> row, one kernel call per (row, occ, icc)
#pragma omp parallel
    {
      const int ithr = omp_get_thread_num(), nthr = omp_get_num_threads();
      int start = 0, end = 0;
      utils::balance211(j.bs * j.oh, nthr, ithr, start, end);
      std::vector<s32> ws(ws_row * j.oh), ws1(ws1_row * j.oh);
      int n = 0, oh_s = 0;
      utils::nd_iterator_init(start, n, j.bs, oh_s, j.oh);
      while (start < end) {
        const int oh_e = std::min(j.oh, oh_s + (end - start));
        for (int occ = 0; occ < oc_chunks; ++occ) {
          const int ocb = occ * j.nb_oc_blocking;
          for (int icc = 0; icc < ic_chunks; ++icc) {
            const int icb = icc * j.nb_ic_blocking;
            for (int oj = oh_s; oj < oh_e; ++oj) {
              const int ij = oj * j.sh - j.t_pad;
              const int t_over = std::max(0, -ij), b_over = std::max(j.ih, ij + j.kh) - j.ih;
              jit::jit_conv_call_t p = {0};
              p.src = src + ((size_t)(n * j.ih + ij + t_over) * j.iw) * j.ic + (size_t)icb * j.ic_block;
              p.wei = wei + (((size_t)ocb * j.nb_ic + icb) * j.kh + t_over) * j.kw * blk;
              p.bia = bias ? bias + (size_t)ocb * j.oc_block * j.typesize_conv0_bia : nullptr;
              p.scales = s0.v.data() + (j.conv0_multi_oc_scale ? ocb * j.oc_block : 0);
              p.acc_s32 = ws.data() + (size_t)(oj - oh_s) * ws_row;
              p.channel = icb;
              p.kh_padding = std::max(0, j.kh - t_over - b_over);
              p.ocb3x3 = ocb;
              p.wei1x1 = wei1 + (size_t)ocb * blk;
              p.bia1x1 = bia1;
              p.scales1x1 = s1.v.data();
              p.acc1x1 = ws1.data() + (size_t)(oj - oh_s) * ws1_row;
              p.dst = out + ((size_t)(n * j.oh + oj) * j.ow) * j.oc1x1 * j.typesize_out;
              kernel.jit_ker_(&p);
            }
          }
        }
        utils::nd_iterator_jump(start, end, n, j.bs, oh_s, j.oh);
      }
    }
  } else {
    // corrected restatement of op_conv<T>::infer_conv0 (src/op_conv.cc:31-138), loop order loop_cgn
#pragma omp parallel
    {
      const int ithr = omp_get_thread_num(), nthr = omp_get_num_threads();
      int start = 0, end = 0;
      utils::balance211(j.bs * oc_chunks * j.oh, nthr, ithr, start, end);
      std::vector<s32> ws(ws_row * j.oh);
      int n = 0, occ = 0, oh_s = 0;
      utils::nd_iterator_init(start, occ, oc_chunks, n, j.bs, oh_s, j.oh);
      while (start < end) {
        const int ocb = occ * j.nb_oc_blocking;
        const int oh_e = std::min(j.oh, oh_s + (end - start));
        for (int icc = 0; icc < ic_chunks; ++icc) {
          const int icb = icc * j.nb_ic_blocking;
          for (int oj = oh_s; oj < oh_e; ++oj) {
            const int ij = oj * j.sh - j.t_pad;
            const int t_over = std::max(0, -ij), b_over = std::max(j.ih, ij + j.kh) - j.ih;
            jit::jit_conv_call_t p = {0};
            p.src = src + ((size_t)(n * j.ih + ij + t_over) * j.iw) * j.ic + (size_t)icb * j.ic_block;
            p.wei = wei + (((size_t)ocb * j.nb_ic + icb) * j.kh + t_over) * j.kw * blk;
            p.bia = bias ? bias + (size_t)ocb * j.oc_block * j.typesize_conv0_bia : nullptr;
            p.scales = s0.v.data() + (j.conv0_multi_oc_scale ? ocb * j.oc_block : 0);
            p.acc_s32 = ws.data() + (size_t)(oj - oh_s) * ws_row;
            p.channel = icb;
            p.kh_padding = std::max(0, j.kh - t_over - b_over);
            p.dst = out + (((size_t)(n * j.oh + oj) * j.ow) * j.oc + (size_t)ocb * j.oc_block) * j.typesize_out;
            kernel.jit_ker_(&p);
          }
        }
        utils::nd_iterator_jump(start, end, occ, oc_chunks, n, j.bs, oh_s, j.oh);
      }
    }
  }
  return 0;
}
