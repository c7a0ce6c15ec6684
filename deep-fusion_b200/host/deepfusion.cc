// deepfusion.cc -- C++ host layer: the reference's public API (include/deepfusion.h) on top of the
// CUDA C-ABI (include/dfcuda.h).
//
// Mirrors, function by function, the reference's L3/L2 layers:
//   memory            src/deepfusion.cc:25-88, util/memory.cc
//   op::submit        src/deepfusion.cc:90-103
//   concat()/conv()   src/deepfusion.cc:105-185
//   op_concat<T>      src/op_concat.{h,cc}          -> concat_op
//   op_conv<T>        src/op_conv.{h,cc}            -> conv_op
// Differences from the reference are deliberate and listed in DESIGN.md ("Deviations"): weights,
// biases and scales are captured (copied to the device) at creation; std_dims() is initialised by
// both constructors; error text is the same, the process still exits on creation failure.
#include <assert.h>
#include <string.h>
#include <sys/time.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "df_host.h"

namespace deepfusion {
namespace detail {

size_t dtype_size(memory::dtype dt) {
  switch (dt) {
    case memory::dtype::f32:
    case memory::dtype::s32: return 4;
    case memory::dtype::s8:
    case memory::dtype::u8: return 1;
    default: assert(!"Unkown data type"); return 0;
  }
}

int conv_output_size(int image, int kernel, int stride, int padding) {
  return (image + 2 * padding - kernel) / stride + 1;
}

bool profiling_enabled() {
  // the reference reads DEEPFUSION_PROFILE (util/scaffold.cc:56-66) while its README documents
  // DEEPFUSION_VERBOSE (README.md:26-31); accept either (defect D6)
  static int cached = -1;
  if (cached < 0) {
    const char *a = getenv("DEEPFUSION_VERBOSE"), *b = getenv("DEEPFUSION_PROFILE");
    cached = ((a && atoi(a) != 0) || (b && atoi(b) != 0)) ? 1 : 0;
  }
  return cached == 1;
}

static double now_ms() {
  struct timeval t;
  gettimeofday(&t, NULL);
  return 1e+3 * t.tv_sec + 1e-3 * t.tv_usec;
}

static void cuda_or_exit(int rc, const char *what) {
  if (rc != 0) error_and_exit("%s failed (%d): %s", what, rc, df_last_error());
}

// include/deepfusion.h is the reference's header, byte for byte: `memory` has no member for the device mirror
// and `op` has no virtual destructor.  What the B200 layer adds to both lives in side tables keyed by the
// object's address.
static std::mutex g_mu;
static std::unordered_map<const memory *, memory_state> &memory_states() {
  static std::unordered_map<const memory *, memory_state> *t = new std::unordered_map<const memory *, memory_state>();
  return *t;  // (never destroyed: memories with static storage may outlive any table destructor)
}
memory_state *state_of(memory &m) {
  std::lock_guard<std::mutex> g(g_mu);
  return &memory_states()[&m];  // node-based map: the address stays valid until erased
}
static void drop_state(memory &m) {
  memory_state st;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = memory_states().find(&m);
    if (it == memory_states().end()) return;
    st = it->second;
    memory_states().erase(it);
  }
  if (st.pinned) df_host_unregister(m.data());
  if (st.dev) df_free(st.dev);
}

// An op created by the factories is destroyed through `std::unique_ptr<op>` -- i.e. through a base class
// WITHOUT a virtual destructor (reference include/deepfusion.h:105-114), so a derived destructor never
// runs.  Everything an op owns therefore sits in an op_resources object in this table; it is released
//   * explicitly by ext::release(op), or
//   * when a new op is created at the same address (the old one can only be gone), or
//   * at process exit.
// The reference's own op_conv / op_concat leak their workspaces in exactly this situation.
static std::unordered_map<const op *, op_resources *> &op_table() {
  static std::unordered_map<const op *, op_resources *> *t = new std::unordered_map<const op *, op_resources *>();
  return *t;
}
void adopt_resources(const op *o, op_resources *r) {
  op_resources *stale = nullptr;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = op_table().find(o);
    if (it != op_table().end()) {
      stale = it->second;
      it->second = r;
    } else {
      op_table()[o] = r;
    }
  }
  delete stale;
}
bool release_resources(const op *o) {
  op_resources *r = nullptr;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = op_table().find(o);
    if (it == op_table().end()) return false;
    r = it->second;
    op_table().erase(it);
  }
  delete r;
  return true;
}

static void *mirror(memory &m) {
  memory_state *st = state_of(m);
  if (!st->dev) cuda_or_exit(df_malloc(m.buffer_size(), &st->dev), "device allocation");
  return st->dev;
}

static int dt_code(memory::dtype dt) { return static_cast<int>(dt); }  // same numbering as DF_*

}  // namespace detail

using detail::cuda_or_exit;
using detail::mirror;

// ------------------------------------------------------------------------------- memory
static memory::dims nchw2format(const memory::nchw_dims &dm, const memory::format fmt) {
  memory::dims out(4);
  switch (fmt) {
    case memory::format::nhwc:
      out[0] = dm[0];
      out[1] = dm[2];
      out[2] = dm[3];
      out[3] = dm[1];
      break;
    case memory::format::nchw:
    case memory::format::OIhw4i16o4i:
      out[0] = dm[0];
      out[1] = dm[1];
      out[2] = dm[2];
      out[3] = dm[3];
      break;
    default: error_and_exit("bad type");
  }
  return out;
}

memory::memory(const nchw_dims &dm, const format fmt, const dtype dt, int alignment)
    : data_(nullptr), std_dims_(dm), fmt_(fmt), dt_(dt) {
  dims_ = nchw2format(dm, fmt);
  allocate_buffer(alignment);
}

memory::memory(const dims &dm, const format fmt, const dtype dt, int alignment)
    : data_(nullptr), dims_(dm), fmt_(fmt), dt_(dt) {
  // the reference leaves std_dims_ uninitialised here although op_conv reads it (defect D5)
  for (size_t i = 0; i < 4; ++i) std_dims_[i] = i < dm.size() ? dm[i] : 1;
  allocate_buffer(alignment);
}

memory::~memory() {
  detail::drop_state(*this);
  free(data_);
}

void memory::allocate_buffer(int alignment) {
  assert(buffer_size() > 0);
  void *p = nullptr;
  if (::posix_memalign(&p, alignment, buffer_size()) != 0) p = nullptr;
  data_ = p;
  assert(data_ != NULL);
}

size_t memory::size() {
  size_t n = 1;
  for (size_t i = 0; i < dims_.size(); ++i) n *= size_t(dims_[i]);
  return n;
}

size_t memory::buffer_size() { return size() * detail::dtype_size(dt_); }

// ----------------------------------------------------------------------------------- op
void op::submit() {
  double t0 = 0;
  const bool prof = detail::profiling_enabled();
  if (prof) t0 = detail::now_ms();
  infer();
  if (prof) info("%s infer %f", this->name(), detail::now_ms() - t0);
}

// ------------------------------------------------------------------------------- concat
namespace {

class concat_op : public detail::device_op {
public:
  concat_op(const std::vector<std::unique_ptr<memory>> &srcs, std::unique_ptr<memory> &dst, bool post_relu)
      : relu_(post_relu), dst_(dst.get()), res_(new resources()), srcs_(res_->srcs), ic_(res_->ic) {
    detail::adopt_resources(this, res_);
    if (!init_conf(srcs, dst)) error_and_exit("Init Concat op failed!");
    for (size_t i = 0; i < srcs.size(); ++i) srcs_.push_back(srcs[i].get());
  }

  void launch(void *stream) override {
    std::vector<const void *> ptrs(srcs_.size());
    for (size_t i = 0; i < srcs_.size(); ++i) ptrs[i] = mirror(*srcs_[i]);
    cuda_or_exit(df_concat_run(detail::dt_code(dst_->data_type()), relu_, (int)srcs_.size(), ptrs.data(), ic_.data(),
                               mirror(*dst_), n_pixels_, stream),
                 "concat launch");
  }
  int launches() const override { return (int)((srcs_.size() + 15) / 16); }

protected:
  // jit_concat_kernel::init_conf (src/jit_concat_kernel.cc:130-197)
  bool init_conf(const std::vector<std::unique_ptr<memory>> &srcs, const std::unique_ptr<memory> &dst) {
    if (srcs.empty() || !dst) return false;
    if (dst->dim_format() != memory::format::nhwc) return false;  // only nhwc
    auto dm = dst->actual_dims();
    if (dm.size() != 4) return false;
    long oc = 0;
    for (size_t i = 0; i < srcs.size(); ++i) {
      if (srcs[i]->dim_format() != dst->dim_format()) return false;
      if (srcs[i]->data_type() != dst->data_type()) return false;
      auto sd = srcs[i]->actual_dims();
      if (sd.size() != 4 || sd[0] != dm[0] || sd[1] != dm[1] || sd[2] != dm[2]) {
        info("Concat input %zu spatial dims do not match", i);
        return false;
      }
      ic_.push_back(sd[3]);
      oc += sd[3];
    }
    if (oc != dm[3]) {
      info("Concat output channels do not match the inputs");
      return false;
    }
    if (df_concat_check(detail::dt_code(dst->data_type()), (int)ic_.size(), ic_.data()) != 0) {
      info("%s", df_last_error());
      return false;
    }
    n_pixels_ = (long)dm[0] * dm[1] * dm[2];
    return true;
  }

  void infer() override {
    for (memory *s : srcs_) cuda_or_exit(df_h2d(mirror(*s), s->data(), s->buffer_size(), nullptr), "concat H2D");
    launch(nullptr);
    cuda_or_exit(df_d2h(dst_->data(), mirror(*dst_), dst_->buffer_size(), nullptr), "concat D2H");
    cuda_or_exit(df_stream_sync(nullptr), "concat sync");
  }
  const char *name() override { return "concat"; }

private:
  struct resources : detail::op_resources {  // heap state of the op (see detail::adopt_resources)
    std::vector<memory *> srcs;
    std::vector<int> ic;
  };
  bool relu_;
  memory *dst_;
  resources *res_;
  std::vector<memory *> &srcs_;
  std::vector<int> &ic_;
  long n_pixels_ = 0;
};

// --------------------------------------------------------------------------------- conv
class conv_op : public detail::device_op {
public:
  conv_op(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
          std::array<int, 2> sz_stride, std::array<int, 2> sz_padding, std::unique_ptr<memory> &dst,
          const std::vector<float> &conv0_scales, const std::vector<float> &conv1_scales,
          const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1, bool conv0_relu,
          bool conv1_relu, round_mode conv0_round_mode, round_mode conv1_round_mode, bool create_handle = true)
      : src_(src.get()), dst_(dst.get()) {
    memset(&desc_, 0, sizeof desc_);
    if (!init_conf(src, wei, bia, sz_stride, sz_padding, dst, conv0_scales, conv1_scales, wei1x1, bia1x1, conv0_relu,
                   conv1_relu, conv0_round_mode, conv1_round_mode))
      error_and_exit("Init Conv op failed!");
    if (!create_handle) return;  // sharded_conv_op creates one handle per device
    // parameters are captured now (the reference keeps raw pointers and a dangling scales
    // pointer, defect D7); the activations stay borrowed
    int rc = df_conv_create(&desc_, static_cast<const int8_t *>(wei->data()),
                            wei1x1 ? static_cast<const int8_t *>(wei1x1->data()) : nullptr,
                            bia ? bia->data() : nullptr, bia1x1 ? bia1x1->data() : nullptr, conv0_scales.data(),
                            conv1_scales.data(), &handle_);
    if (rc == DF_E_UNSUPPORTED) error_and_exit("unsupported on B200 path: %s", df_last_error());
    if (rc != 0) {
      info("%s", df_last_error());
      error_and_exit("Init Conv op failed!");
    }
    res_ = new resources();
    res_->handle = handle_;
    detail::adopt_resources(this, res_);
  }

  void launch(void *stream) override {
    cuda_or_exit(df_conv_run(handle_, static_cast<const uint8_t *>(mirror(*src_)), mirror(*dst_), desc_.n, stream),
                 "conv launch");
  }
  static int submit_slabs(int n) { return n >= 32 ? 8 : (n >= 8 ? 4 : 1); }
  int launches() const override { return submit_slabs(desc_.n); }  // slabs of the pipelined submit()

protected:
  // op_conv<T>::init_conf (src/op_conv.cc:262-365) + the format gate of
  // jit_conv_kernel::init_conf (src/jit_conv_kernel.cc:531-564); the numeric rules are
  // re-checked by df_conv_create.
  bool init_conf(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                 const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                 std::unique_ptr<memory> &dst, const std::vector<float> &conv0_scales,
                 const std::vector<float> &conv1_scales, const std::unique_ptr<memory> &wei1x1,
                 const std::unique_ptr<memory> &bia1x1, bool conv0_relu, bool conv1_relu, round_mode r0,
                 round_mode r1) {
    if (!src || !wei || !dst) return false;
    using dt = memory::dtype;
    using fm = memory::format;
    if (src->data_type() != dt::u8 || wei->data_type() != dt::s8 || (wei1x1 && wei1x1->data_type() != dt::s8)) {
      info("Conv needs u8 src and s8 weights");
      return false;
    }
    auto blocked = [](const std::unique_ptr<memory> &w) {
      return w->dim_format() == fm::OIhw4i16o4i || w->dim_format() == fm::gOIhw4i16o4i;
    };
    if (src->dim_format() != fm::nhwc || dst->dim_format() != fm::nhwc || !blocked(wei) ||
        (wei1x1 && !blocked(wei1x1)) || (bia && bia->dim_format() != fm::x) ||
        (bia1x1 && bia1x1->dim_format() != fm::x)) {
      info("Conv formats must be nhwc / OIhw4i16o4i / x");
      return false;
    }
    constexpr int C = 1, H = 2, W = 3;
    auto src_dims = src->std_dims(), wei_dims = wei->std_dims(), dst_dims = dst->std_dims();
    for (size_t i = 0; i < 2; ++i)
      if (dst_dims[i + 2] != detail::conv_output_size(src_dims[i + 2], wei_dims[i + 2], sz_stride[i], sz_padding[i])) {
        info("Output image size do not match: %zu", i);
        return false;
      }
    if (src_dims[0] != dst_dims[0]) {
      info("Batch size do not equal");
      return false;
    }
    if (src_dims[C] != wei_dims[C]) {
      info("Input channel do not match");
      return false;
    }
    if (!wei1x1) {
      if (dst_dims[C] != wei_dims[0]) {
        info("Output channel do not match");
        return false;
      }
      if (bia && bia->std_dims()[0] != wei_dims[0]) {
        info("Bias channel do not match");
        return false;
      }
      if (conv0_scales.size() != 1 && conv0_scales.size() != size_t(dst_dims[C])) return false;
    } else {
      auto w1 = wei1x1->std_dims();
      if (w1[C] != wei_dims[0]) {
        info("Conv0 output channel do not match");
        return false;
      }
      if (dst_dims[C] != w1[0]) {
        info("Conv1x1 output channel do not match");
        return false;
      }
      if (w1[H] != 1 || w1[W] != 1) {
        info("Fused conv must be 1x1 kernel");
        return false;
      }
      if (bia && bia->std_dims()[0] != wei_dims[0]) {
        info("Bias channel do not match");
        return false;
      }
      if (bia1x1 && bia1x1->std_dims()[0] != dst_dims[C]) {
        info("Bias channel do not match");
        return false;
      }
      if ((conv0_scales.size() != 1 && conv0_scales.size() != size_t(w1[1])) ||
          (conv1_scales.size() != 1 && conv1_scales.size() != size_t(w1[0])))
        return false;
    }
    desc_.n = src_dims[0];
    desc_.ih = src_dims[H];
    desc_.iw = src_dims[W];
    desc_.ic = src_dims[C];
    desc_.oc = wei_dims[0];  // defect D1: the reference takes this from dst
    desc_.oc1 = wei1x1 ? wei1x1->std_dims()[0] : 0;
    desc_.kh = wei_dims[H];
    desc_.kw = wei_dims[W];
    desc_.sh = sz_stride[0];
    desc_.sw = sz_stride[1];
    desc_.ph = sz_padding[0];
    desc_.pw = sz_padding[1];
    desc_.dst_dt = detail::dt_code(dst->data_type());
    desc_.bia0_dt = bia ? detail::dt_code(bia->data_type()) : DF_UNDEF;
    desc_.bia1_dt = bia1x1 ? detail::dt_code(bia1x1->data_type()) : DF_UNDEF;
    desc_.relu0 = conv0_relu;
    desc_.relu1 = conv1_relu;
    desc_.round0 = r0 == round_mode::down ? DF_ROUND_DOWN : DF_ROUND_NEAREST;
    desc_.round1 = r1 == round_mode::down ? DF_ROUND_DOWN : DF_ROUND_NEAREST;
    desc_.nscale0 = (int)conv0_scales.size();
    desc_.nscale1 = (int)conv1_scales.size();
    return true;
  }

  // submit(): host -> device, kernel, device -> host, synchronous like the reference.  Images are
  // independent, so the batch is cut into slabs that flow through three streams: the upload of slab
  // i+1, the kernel of slab i and the download of slab i-1 overlap (PCIe is full duplex), which is
  // what bounds this path -- the kernel itself is ~4 % of it.
  void infer() override {
    const int n = desc_.n;
    const int slabs = submit_slabs(n);
    uint8_t *d_src = static_cast<uint8_t *>(mirror(*src_)), *d_dst = static_cast<uint8_t *>(mirror(*dst_));
    const uint8_t *h_src = static_cast<const uint8_t *>(src_->data());
    uint8_t *h_dst = static_cast<uint8_t *>(dst_->data());
    if (slabs == 1) {
      cuda_or_exit(df_h2d(d_src, h_src, src_->buffer_size(), nullptr), "conv H2D");
      launch(nullptr);
      cuda_or_exit(df_d2h(h_dst, d_dst, dst_->buffer_size(), nullptr), "conv D2H");
      cuda_or_exit(df_stream_sync(nullptr), "conv sync");
      return;
    }
    std::vector<void *> &streams_ = res_->streams, &events_ = res_->events;
    if (streams_.empty()) {
      detail::memory_state *ss = detail::state_of(*src_), *ds = detail::state_of(*dst_);  // pinned buffers make the copies truly asynchronous
      if (!ss->pinned && df_host_register(src_->data(), src_->buffer_size()) == 0) ss->pinned = true;
      if (!ds->pinned && df_host_register(dst_->data(), dst_->buffer_size()) == 0) ds->pinned = true;
      streams_.resize(3);
      for (void *&st : streams_) cuda_or_exit(df_stream_create(&st), "stream create");
      events_.resize(2 * slabs + 2);
      for (void *&e : events_) cuda_or_exit(df_event_create(&e), "event create");
    }
    // The pipeline is the same every time (the op's memories are fixed for its lifetime, src/op_conv.h:82-95), so
    // it is captured ONCE into a CUDA graph -- copies, kernels and the cross-stream dependencies -- and every
    // submit() is one graph launch instead of ~7 runtime calls per slab (their host cost was a fifth of the step).
    if (!res_->graph && !res_->graph_failed) {
      if (df_graph_begin(streams_[0]) == 0) {
        enqueue_pipeline(slabs);
        // join the kernel and download streams back into the capture's origin stream
        cuda_or_exit(df_event_record(events_[2 * slabs], streams_[1]), "event record");
        cuda_or_exit(df_stream_wait_event(streams_[0], events_[2 * slabs]), "stream wait");
        cuda_or_exit(df_event_record(events_[2 * slabs + 1], streams_[2]), "event record");
        cuda_or_exit(df_stream_wait_event(streams_[0], events_[2 * slabs + 1]), "stream wait");
        if (df_graph_end(streams_[0], &res_->graph) != 0) res_->graph_failed = true;
      } else {
        res_->graph_failed = true;
      }
    }
    if (res_->graph) {
      cuda_or_exit(df_graph_launch(res_->graph, streams_[0]), "conv graph launch");
      cuda_or_exit(df_stream_sync(streams_[0]), "conv sync");
      return;
    }
    enqueue_pipeline(slabs);
    cuda_or_exit(df_stream_sync(streams_[2]), "conv sync");
  }

  // upload of slab i+1 | kernel of slab i | download of slab i-1 on three streams
  void enqueue_pipeline(int slabs) {
    const int n = desc_.n;
    uint8_t *d_src = static_cast<uint8_t *>(mirror(*src_)), *d_dst = static_cast<uint8_t *>(mirror(*dst_));
    const uint8_t *h_src = static_cast<const uint8_t *>(src_->data());
    uint8_t *h_dst = static_cast<uint8_t *>(dst_->data());
    const size_t src_img = src_->buffer_size() / n, dst_img = dst_->buffer_size() / n;
    // slab sizes: a small first slab (the downloads -- the long pole, 4x the upload bytes for u8 in / u8 out at
    // OC1 = 4 IC -- start as early as possible), the rest in equal parts
    static const char *plan_env = getenv("DF_SUBMIT_PLAN");  // development knob: "e" = equal slabs
    const bool progressive = !(plan_env && plan_env[0] == 'e') && slabs >= 4 && n >= 4 * slabs;
    const int first_cnt = progressive ? std::max(1, n / (4 * slabs)) : 0;
    const int per = progressive ? (n - first_cnt + slabs - 2) / (slabs - 1) : (n + slabs - 1) / slabs;
    for (int i = 0, first = 0; i < slabs && first < n; ++i) {
      const int want = (progressive && i == 0) ? first_cnt : per;
      const int cnt = first + want <= n ? want : n - first;
      enqueue_slab(i, first, cnt, d_src, d_dst, h_src, h_dst, src_img, dst_img);
      first += cnt;
    }
  }
  void enqueue_slab(int i, int first, int cnt, uint8_t *d_src, uint8_t *d_dst, const uint8_t *h_src, uint8_t *h_dst,
                    size_t src_img, size_t dst_img) {
    std::vector<void *> &streams_ = res_->streams, &events_ = res_->events;
    {
      cuda_or_exit(df_h2d(d_src + first * src_img, h_src + first * src_img, cnt * src_img, streams_[0]), "conv H2D");
      cuda_or_exit(df_event_record(events_[2 * i], streams_[0]), "event record");
      cuda_or_exit(df_stream_wait_event(streams_[1], events_[2 * i]), "stream wait");
      cuda_or_exit(df_conv_run(handle_, d_src + first * src_img, d_dst + first * dst_img, cnt, streams_[1]), "conv launch");
      cuda_or_exit(df_event_record(events_[2 * i + 1], streams_[1]), "event record");
      cuda_or_exit(df_stream_wait_event(streams_[2], events_[2 * i + 1]), "stream wait");
      cuda_or_exit(df_d2h(h_dst + first * dst_img, d_dst + first * dst_img, cnt * dst_img, streams_[2]), "conv D2H");
    }
  }
  const char *name() override { return "conv"; }

private:
  // what the op owns (see detail::adopt_resources: `op` has no virtual destructor)
  struct resources : detail::op_resources {
    df_conv *handle = nullptr;
    std::vector<void *> streams, events;  // created on the first pipelined submit()
    void *graph = nullptr;                // the captured pipeline (see infer())
    bool graph_failed = false;
    ~resources() override {
      if (graph) df_graph_destroy(graph);
      df_conv_destroy(handle);
      for (void *e : events) df_event_destroy(e);
      for (void *st : streams) df_stream_destroy(st);
    }
  };
protected:
  memory *src_, *dst_;
  df_conv_desc desc_;

private:
  df_conv *handle_ = nullptr;
  resources *res_ = nullptr;
};

// ------------------------------------------------------------------------ conv, batch-sharded
// The same operator over the GPUs of one box: every image is independent -- the reference itself
// splits the work over n * oh rows with balance211 (src/op_conv.cc:155-156) -- so device g owns the contiguous
// slab of ceil(N / G) images starting at g * ceil(N / G); NHWC makes that one byte range of src and dst.
// Weights / biases / scales are replicated at creation, there is no data-path collective and no NCCL.  ONE host
// thread drives every device: per device three streams (upload, kernel, download) and sub-slabs, so that the
// PCIe directions and the kernels of all devices overlap; submit() returns when the last download has landed.
class sharded_conv_op : public conv_op {
public:
  sharded_conv_op(const std::vector<int> &devices, const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                  const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                  std::unique_ptr<memory> &dst, const std::vector<float> &conv0_scales, const std::vector<float> &conv1_scales,
                  const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1, bool conv0_relu, bool conv1_relu,
                  round_mode r0, round_mode r1)
      : conv_op(src, wei, bia, sz_stride, sz_padding, dst, conv0_scales, conv1_scales, wei1x1, bia1x1, conv0_relu, conv1_relu,
                r0, r1, /*create_handle=*/false) {
    if (devices.empty()) error_and_exit("conv_sharded: no devices");
    sh_ = new shards();
    detail::adopt_resources(this, sh_);
    int count = 0;
    cuda_or_exit(df_get_device(&home_), "get device");
    cuda_or_exit(df_device_count(&count), "device count");
    for (int d : devices)
      if (d < 0 || d >= count) error_and_exit("conv_sharded: device %d does not exist (%d devices)", d, count);
    const int G = (int)devices.size(), n = desc_.n, per = (n + G - 1) / G;
    detail::memory_state *ss = detail::state_of(*src_), *ds = detail::state_of(*dst_);
    if (!ss->pinned && df_host_register(src_->data(), src_->buffer_size()) == 0) ss->pinned = true;
    if (!ds->pinned && df_host_register(dst_->data(), dst_->buffer_size()) == 0) ds->pinned = true;
    const size_t src_img = src_->buffer_size() / n, dst_img = dst_->buffer_size() / n;
    for (int g = 0; g < G; ++g) {
      shard s;
      s.device = devices[g];
      s.first = std::min(g * per, n);
      s.count = std::max(0, std::min(per, n - s.first));
      if (s.count > 0) {
        cuda_or_exit(df_set_device(s.device), "set device");
        df_conv_desc d = desc_;
        d.n = s.count;
        int rc = df_conv_create(&d, static_cast<const int8_t *>(wei->data()), wei1x1 ? static_cast<const int8_t *>(wei1x1->data()) : nullptr,
                                bia ? bia->data() : nullptr, bia1x1 ? bia1x1->data() : nullptr, conv0_scales.data(),
                                conv1_scales.data(), &s.handle);
        if (rc == DF_E_UNSUPPORTED) error_and_exit("unsupported on B200 path: %s", df_last_error());
        if (rc != 0) {
          info("%s", df_last_error());
          error_and_exit("Init Conv op failed!");
        }
        cuda_or_exit(df_malloc(s.count * src_img, &s.d_src), "device allocation");
        cuda_or_exit(df_malloc(s.count * dst_img, &s.d_dst), "device allocation");
        for (void *&st : s.streams) cuda_or_exit(df_stream_create(&st), "stream create");
        for (void *&e : s.events) cuda_or_exit(df_event_create(&e), "event create");
      }
      sh_->v.push_back(s);
    }
    cuda_or_exit(df_set_device(home_), "set device");
  }

  // kernels only, on slabs already resident on the devices (ext::submit_device): asynchronous
  void launch(void *) override {
    for (shard &s : sh_->v) {
      if (!s.count) continue;
      cuda_or_exit(df_set_device(s.device), "set device");
      cuda_or_exit(df_conv_run(s.handle, static_cast<const uint8_t *>(s.d_src), s.d_dst, s.count, s.streams[1]), "conv launch");
    }
    cuda_or_exit(df_set_device(home_), "set device");
  }
  int launches() const override {
    int k = 0;
    for (const shard &s : sh_->v) k += s.count > 0 ? (s.count >= 8 ? kSub : 1) : 0;
    return k;
  }
  void sync_all() {
    for (shard &s : sh_->v) {
      if (!s.count) continue;
      cuda_or_exit(df_set_device(s.device), "set device");
      cuda_or_exit(df_stream_sync(s.streams[1]), "sync");
      cuda_or_exit(df_stream_sync(s.streams[2]), "sync");
    }
    cuda_or_exit(df_set_device(home_), "set device");  // the caller's device, as at creation
  }
  void upload() {  // host src -> every device's slab (for device-resident timing)
    const size_t src_img = src_->buffer_size() / desc_.n;
    for (shard &s : sh_->v) {
      if (!s.count) continue;
      cuda_or_exit(df_set_device(s.device), "set device");
      cuda_or_exit(df_h2d(s.d_src, static_cast<const uint8_t *>(src_->data()) + s.first * src_img, s.count * src_img, s.streams[1]), "H2D");
    }
    sync_all();
  }
  void download() {
    const size_t dst_img = dst_->buffer_size() / desc_.n;
    for (shard &s : sh_->v) {
      if (!s.count) continue;
      cuda_or_exit(df_set_device(s.device), "set device");
      cuda_or_exit(df_d2h(static_cast<uint8_t *>(dst_->data()) + s.first * dst_img, s.d_dst, s.count * dst_img, s.streams[1]), "D2H");
    }
    sync_all();
  }
  int device_count_used() const {
    int k = 0;
    for (const shard &s : sh_->v) k += s.count > 0;
    return k;
  }

protected:
  void infer() override {
    const int n = desc_.n;
    const size_t src_img = src_->buffer_size() / n, dst_img = dst_->buffer_size() / n;
    const uint8_t *h_src = static_cast<const uint8_t *>(src_->data());
    uint8_t *h_dst = static_cast<uint8_t *>(dst_->data());
    // sub-slab k of every device before sub-slab k + 1 of any: all devices start uploading at once
    for (int k = 0; k < kSub; ++k)
      for (shard &s : sh_->v) {
        if (!s.count) continue;
        const int subs = s.count >= 8 ? kSub : 1;
        if (k >= subs) continue;
        const int per = (s.count + subs - 1) / subs, first = k * per;
        const int cnt = first + per <= s.count ? per : s.count - first;
        if (cnt <= 0) continue;
        cuda_or_exit(df_set_device(s.device), "set device");
        uint8_t *d_src = static_cast<uint8_t *>(s.d_src) + first * src_img, *d_dst = static_cast<uint8_t *>(s.d_dst) + first * dst_img;
        cuda_or_exit(df_h2d(d_src, h_src + (s.first + first) * src_img, cnt * src_img, s.streams[0]), "conv H2D");
        cuda_or_exit(df_event_record(s.events[2 * k], s.streams[0]), "event record");
        cuda_or_exit(df_stream_wait_event(s.streams[1], s.events[2 * k]), "stream wait");
        cuda_or_exit(df_conv_run(s.handle, d_src, d_dst, cnt, s.streams[1]), "conv launch");
        cuda_or_exit(df_event_record(s.events[2 * k + 1], s.streams[1]), "event record");
        cuda_or_exit(df_stream_wait_event(s.streams[2], s.events[2 * k + 1]), "stream wait");
        cuda_or_exit(df_d2h(h_dst + (s.first + first) * dst_img, d_dst, cnt * dst_img, s.streams[2]), "conv D2H");
      }
    sync_all();
  }
  const char *name() override { return "conv_sharded"; }

private:
  static constexpr int kSub = 4;
  struct shard {
    int device = 0, first = 0, count = 0;
    df_conv *handle = nullptr;
    void *d_src = nullptr, *d_dst = nullptr;
    void *streams[3] = {nullptr, nullptr, nullptr};
    void *events[2 * kSub] = {};
  };
  struct shards : detail::op_resources {
    std::vector<shard> v;
    ~shards() override {
      for (shard &s : v) {
        if (!s.count) continue;
        df_set_device(s.device);
        df_conv_destroy(s.handle);
        df_free(s.d_src);
        df_free(s.d_dst);
        for (void *e : s.events) df_event_destroy(e);
        for (void *st : s.streams) df_stream_destroy(st);
      }
    }
  };
  shards *sh_ = nullptr;
  int home_ = 0;
};

// ------------------------------------------------------------------ concat fused into the conv
// ext::concat_conv(): the producer -> consumer pair op_concat<u8> (src/op_concat.cc:22-72) -> op_conv<T>
// (src/op_conv.cc:140-260) as ONE operator.  The concatenated tensor is never materialised: the conv kernel's
// halo loads read each K-block from the input that owns those channels (df_conv_create_concat).  Checks = the
// concat's (same dtype / format / N,H,W, jit_concat_kernel.cc:178-190) + the conv's, made against a shape-only
// stand-in for the concatenated source.  When the fused load cannot take the channel split (an input that is a
// multiple of 16 but not of 32) the op runs the two kernels back to back on the device instead.
class concat_conv_op : public conv_op {
public:
  concat_conv_op(std::unique_ptr<memory> &cat_shape, const std::vector<std::unique_ptr<memory>> &srcs, bool concat_relu,
                 const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride,
                 std::array<int, 2> sz_padding, std::unique_ptr<memory> &dst, const std::vector<float> &conv0_scales,
                 const std::vector<float> &conv1_scales, const std::unique_ptr<memory> &wei1x1,
                 const std::unique_ptr<memory> &bia1x1, bool conv0_relu, bool conv1_relu, round_mode r0, round_mode r1)
      : conv_op(cat_shape, wei, bia, sz_stride, sz_padding, dst, conv0_scales, conv1_scales, wei1x1, bia1x1, conv0_relu,
                conv1_relu, r0, r1, /*create_handle=*/false),
        relu_(concat_relu) {
    cc_ = new cc_resources();
    detail::adopt_resources(this, cc_);
    cc_->shape = std::move(cat_shape);
    for (auto &m : srcs) {
      cc_->srcs.push_back(m.get());
      cc_->ic.push_back(m->actual_dims()[3]);
    }
    const int8_t *w0 = static_cast<const int8_t *>(wei->data());
    const int8_t *w1 = wei1x1 ? static_cast<const int8_t *>(wei1x1->data()) : nullptr;
    const void *b0 = bia ? bia->data() : nullptr, *b1 = bia1x1 ? bia1x1->data() : nullptr;
    // Which route: the plain conv of this shape tells.  Where its weights have to stream it runs on CTA pairs, and the
    // fused load (32 / 64-byte halo K-blocks, single-CTA kernel) is slower than concat + that kernel (measured,
    // 48.7 vs 33.3 us on BASELINE configs[1]'s shape) -- the op then runs the two kernels back to back,
    // the concatenated tensor staying in the L2.  DEEPFUSION_CONCAT_FUSE=1 / =0 forces the route.
    df_conv *plain = nullptr;
    int rc = df_conv_create(&desc_, w0, w1, b0, b1, conv0_scales.data(), conv1_scales.data(), &plain);
    if (rc == DF_E_UNSUPPORTED) error_and_exit("unsupported on B200 path: %s", df_last_error());
    if (rc == 0) {
      df_conv_info ci;
      const char *force = getenv("DEEPFUSION_CONCAT_FUSE");
      const bool pair_route = df_conv_query(plain, &ci) == 0 && ci.w0_resident == 3;
      const bool want_fused = force ? atoi(force) != 0 : !pair_route;
      if (want_fused && df_conv_create_concat(&desc_, (int)cc_->ic.size(), cc_->ic.data(), concat_relu, w0, w1, b0, b1,
                                              conv0_scales.data(), conv1_scales.data(), &cc_->handle) == 0) {
        df_conv_destroy(plain);
      } else {  // two kernels, concatenated tensor in a device buffer
        fused_ = false;
        cc_->handle = plain;
        cuda_or_exit(df_malloc(cc_->shape->buffer_size(), &cc_->d_cat), "device allocation");
      }
    }
    if (rc != 0) {
      info("%s", df_last_error());
      error_and_exit("Init Conv op failed!");
    }
  }

  void launch(void *stream) override {
    std::vector<const void *> ptrs(cc_->srcs.size());
    for (size_t i = 0; i < ptrs.size(); ++i) ptrs[i] = mirror(*cc_->srcs[i]);
    if (fused_) {
      cuda_or_exit(df_conv_run_concat(cc_->handle, ptrs.data(), mirror(*dst_), desc_.n, stream), "concat+conv launch");
    } else {
      cuda_or_exit(df_concat_run(DF_U8, relu_, (int)ptrs.size(), ptrs.data(), cc_->ic.data(), cc_->d_cat,
                                 (long)desc_.n * desc_.ih * desc_.iw, stream), "concat launch");
      cuda_or_exit(df_conv_run(cc_->handle, static_cast<const uint8_t *>(cc_->d_cat), mirror(*dst_), desc_.n, stream), "conv launch");
    }
  }
  int launches() const override { return fused_ ? 1 : 2; }
  bool fused() const { return fused_; }

protected:
  void infer() override {
    for (memory *m : cc_->srcs) cuda_or_exit(df_h2d(mirror(*m), m->data(), m->buffer_size(), nullptr), "concat+conv H2D");
    launch(nullptr);
    cuda_or_exit(df_d2h(dst_->data(), mirror(*dst_), dst_->buffer_size(), nullptr), "concat+conv D2H");
    cuda_or_exit(df_stream_sync(nullptr), "concat+conv sync");
  }
  const char *name() override { return "concat+conv"; }

private:
  struct cc_resources : detail::op_resources {
    std::unique_ptr<memory> shape;  // shape-only stand-in for the concatenated source (its buffer is never touched)
    std::vector<memory *> srcs;
    std::vector<int> ic;
    df_conv *handle = nullptr;
    void *d_cat = nullptr;
    ~cc_resources() override {
      df_conv_destroy(handle);
      if (d_cat) df_free(d_cat);
    }
  };
  cc_resources *cc_ = nullptr;
  bool relu_, fused_ = true;
};

// --------------------------------------------------- the reference's planned operators (README.md:64-65)
// conv (+ReLU) + pooling: the planned jit_avx512_core_u8s8s32x_convolution_relu_pool_op
// (test/test_conv_relu_pooling.cc:262-278 sketches its arguments: conv src / wei / bia / stride / padding / dst,
// scales, relu, round mode, then pool dst / stride / padding / kernel / round mode).  The conv stage is the conv-only
// operator writing `conv_dst`'s device mirror; the pooling kernel follows on the same stream and reads it from L2.
// submit() uploads the source and downloads only the pooled result.
class conv_pool_op : public conv_op {
public:
  conv_pool_op(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
               std::array<int, 2> sz_stride, std::array<int, 2> sz_padding, std::unique_ptr<memory> &conv_dst,
               std::unique_ptr<memory> &pool_dst, int kind, std::array<int, 2> pool_kernel, std::array<int, 2> pool_stride,
               std::array<int, 2> pool_padding, bool conv_relu, const std::vector<float> &conv_scales, round_mode conv_round,
               round_mode pool_round)
      : conv_op(src, wei, bia, sz_stride, sz_padding, conv_dst, conv_scales, std::vector<float>{1.f}, nullptr, nullptr,
                conv_relu, false, conv_round, round_mode::nearest),
        pool_dst_(pool_dst.get()) {
    if (!pool_dst || pool_dst->dim_format() != memory::format::nhwc || pool_dst->data_type() != conv_dst->data_type()) {
      info("Pooling destination must be nhwc with the conv destination's data type");
      error_and_exit("Init Pooling op failed!");
    }
    const memory::dims c = conv_dst->actual_dims(), o = pool_dst->actual_dims();  // nhwc
    memset(&pd_, 0, sizeof pd_);
    pd_.dtype = detail::dt_code(conv_dst->data_type());
    pd_.kind = kind;
    pd_.n = c[0];
    pd_.h = c[1];
    pd_.w = c[2];
    pd_.c = c[3];
    pd_.kh = pool_kernel[0];
    pd_.kw = pool_kernel[1];
    pd_.sh = pool_stride[0];
    pd_.sw = pool_stride[1];
    pd_.ph = pool_padding[0];
    pd_.pw = pool_padding[1];
    pd_.oh = o.size() == 4 ? o[1] : 0;
    pd_.ow = o.size() == 4 ? o[2] : 0;
    pd_.round_mode = pool_round == round_mode::down ? DF_ROUND_DOWN : DF_ROUND_NEAREST;
    // output size: floor((h + 2p - k) / s) + 1, or one more where a last window still starts inside the image (padR)
    const int oh0 = (pd_.h + 2 * pd_.ph - pd_.kh) / pd_.sh + 1, ow0 = (pd_.w + 2 * pd_.pw - pd_.kw) / pd_.sw + 1;
    if (o.size() != 4 || o[0] != c[0] || o[3] != c[3] || (pd_.oh != oh0 && pd_.oh != oh0 + 1) || (pd_.ow != ow0 && pd_.ow != ow0 + 1) ||
        df_pool_check(&pd_) != 0) {
      info("Pooling output size do not match: %s", df_last_error());
      error_and_exit("Init Pooling op failed!");
    }
  }
  void launch(void *stream) override {
    conv_op::launch(stream);
    cuda_or_exit(df_pool_run(&pd_, mirror(*dst_), mirror(*pool_dst_), pd_.n, stream), "pool launch");
  }
  int launches() const override { return 2; }

protected:
  void infer() override {
    cuda_or_exit(df_h2d(mirror(*src_), src_->data(), src_->buffer_size(), nullptr), "conv H2D");
    launch(nullptr);
    cuda_or_exit(df_d2h(pool_dst_->data(), mirror(*pool_dst_), pool_dst_->buffer_size(), nullptr), "pool D2H");
    cuda_or_exit(df_stream_sync(nullptr), "conv+pool sync");
  }
  const char *name() override { return "conv+relu+pooling"; }

private:
  memory *pool_dst_;
  df_pool_desc pd_;
};

// the pooling stage alone (df_pool_run) as an operator of its own: what follows a conv_sum in the reference's ResNet entry
// (test/test_conv_relu_pooling.cc:341-342: 1x1 conv + eltwise sum + ReLU + 7x7 average pooling)
class pool_op : public detail::device_op {
public:
  pool_op(const std::unique_ptr<memory> &src, std::unique_ptr<memory> &dst, int kind, std::array<int, 2> k, std::array<int, 2> st,
          std::array<int, 2> pad, round_mode rm)
      : src_(src.get()), dst_(dst.get()) {
    if (!src || !dst || src->dim_format() != memory::format::nhwc || dst->dim_format() != memory::format::nhwc ||
        src->data_type() != dst->data_type()) {
      info("Pooling needs nhwc source and destination of one data type");
      error_and_exit("Init Pooling op failed!");
    }
    const memory::dims c = src->actual_dims(), o = dst->actual_dims();
    memset(&pd_, 0, sizeof pd_);
    pd_.dtype = detail::dt_code(src->data_type());
    pd_.kind = kind;
    pd_.n = c[0]; pd_.h = c[1]; pd_.w = c[2]; pd_.c = c[3];
    pd_.kh = k[0]; pd_.kw = k[1]; pd_.sh = st[0]; pd_.sw = st[1]; pd_.ph = pad[0]; pd_.pw = pad[1];
    pd_.oh = o.size() == 4 ? o[1] : 0;
    pd_.ow = o.size() == 4 ? o[2] : 0;
    pd_.round_mode = rm == round_mode::down ? DF_ROUND_DOWN : DF_ROUND_NEAREST;
    if (o.size() != 4 || o[0] != c[0] || o[3] != c[3] || df_pool_check(&pd_) != 0) {
      info("Pooling output size do not match: %s", df_last_error());
      error_and_exit("Init Pooling op failed!");
    }
  }
  void launch(void *stream) override {
    cuda_or_exit(df_pool_run(&pd_, mirror(*src_), mirror(*dst_), pd_.n, stream), "pool launch");
  }
  int launches() const override { return 1; }

protected:
  void infer() override {
    cuda_or_exit(df_h2d(mirror(*src_), src_->data(), src_->buffer_size(), nullptr), "pool H2D");
    launch(nullptr);
    cuda_or_exit(df_d2h(dst_->data(), mirror(*dst_), dst_->buffer_size(), nullptr), "pool D2H");
    cuda_or_exit(df_stream_sync(nullptr), "pool sync");
  }
  const char *name() override { return "pooling"; }

private:
  memory *src_, *dst_;
  df_pool_desc pd_;
};

// conv / fused conv + eltwise sum + ReLU (README.md:65): `residual` has dst's dims and type and is added to the scaled
// result before the ReLU (df_conv_create_sum / df_conv_run_sum).
class conv_sum_op : public conv_op {
public:
  conv_sum_op(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
              std::array<int, 2> sz_stride, std::array<int, 2> sz_padding, const std::unique_ptr<memory> &wei1x1,
              const std::unique_ptr<memory> &bia1x1, const std::unique_ptr<memory> &residual, std::unique_ptr<memory> &dst,
              bool conv0_relu, const std::vector<float> &conv0_scales, round_mode r0, bool conv1_relu,
              const std::vector<float> &conv1_scales, round_mode r1)
      : conv_op(src, wei, bia, sz_stride, sz_padding, dst, conv0_scales, conv1_scales, wei1x1, bia1x1, conv0_relu, conv1_relu,
                r0, r1, /*create_handle=*/false),
        res_mem_(residual.get()) {
    if (!residual || residual->data_type() != dst->data_type() || residual->dim_format() != memory::format::nhwc ||
        residual->actual_dims() != dst->actual_dims()) {
      info("Eltwise-sum input must have the destination's dims, format and data type");
      error_and_exit("Init Conv op failed!");
    }
    sr_ = new sum_resources();
    detail::adopt_resources(this, sr_);
    int rc = df_conv_create_sum(&desc_, static_cast<const int8_t *>(wei->data()),
                                wei1x1 ? static_cast<const int8_t *>(wei1x1->data()) : nullptr, bia ? bia->data() : nullptr,
                                bia1x1 ? bia1x1->data() : nullptr, conv0_scales.data(), conv1_scales.data(), &sr_->handle);
    if (rc == DF_E_UNSUPPORTED) error_and_exit("unsupported on B200 path: %s", df_last_error());
    if (rc != 0) {
      info("%s", df_last_error());
      error_and_exit("Init Conv op failed!");
    }
  }
  void launch(void *stream) override {
    cuda_or_exit(df_conv_run_sum(sr_->handle, static_cast<const uint8_t *>(mirror(*src_)), mirror(*res_mem_), mirror(*dst_),
                                 desc_.n, stream), "conv+sum launch");
  }
  int launches() const override { return 1; }

protected:
  void infer() override {
    cuda_or_exit(df_h2d(mirror(*src_), src_->data(), src_->buffer_size(), nullptr), "conv H2D");
    cuda_or_exit(df_h2d(mirror(*res_mem_), res_mem_->data(), res_mem_->buffer_size(), nullptr), "residual H2D");
    launch(nullptr);
    cuda_or_exit(df_d2h(dst_->data(), mirror(*dst_), dst_->buffer_size(), nullptr), "conv D2H");
    cuda_or_exit(df_stream_sync(nullptr), "conv+sum sync");
  }
  const char *name() override { return "conv+eltwise-sum+relu"; }

private:
  struct sum_resources : detail::op_resources {
    df_conv *handle = nullptr;
    ~sum_resources() override { df_conv_destroy(handle); }
  };
  memory *res_mem_;
  sum_resources *sr_ = nullptr;
};

}  // namespace

// ---------------------------------------------------------------------------- factories
std::unique_ptr<op> concat(const std::vector<std::unique_ptr<memory>> &srcs, std::unique_ptr<memory> &dst,
                           bool post_relu) {
  switch (dst->data_type()) {
    case memory::dtype::f32:
    case memory::dtype::s32:
    case memory::dtype::s8:
    case memory::dtype::u8: return std::unique_ptr<op>(new concat_op(srcs, dst, post_relu));
    default: assert(!"bad data_type");
  }
  return nullptr;
}

std::unique_ptr<op> conv(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                         const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride,
                         std::array<int, 2> sz_padding, const std::unique_ptr<memory> &wei1x1,
                         const std::unique_ptr<memory> &bia1x1, std::unique_ptr<memory> &dst, bool conv0_relu,
                         std::vector<float> conv0_scales, round_mode conv0_round_mode, bool conv1_relu,
                         std::vector<float> conv1_scales, round_mode conv1_round_mode) {
  switch (dst->data_type()) {
    case memory::dtype::f32:
    case memory::dtype::s32:
    case memory::dtype::s8:
    case memory::dtype::u8:
      return std::unique_ptr<op>(new conv_op(src, wei, bia, sz_stride, sz_padding, dst, conv0_scales, conv1_scales,
                                             wei1x1, bia1x1, conv0_relu, conv1_relu, conv0_round_mode,
                                             conv1_round_mode));
    default: assert(!"bad data_type");
  }
  return nullptr;
}

std::unique_ptr<op> conv(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                         const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride,
                         std::array<int, 2> sz_padding, std::unique_ptr<memory> &dst, bool conv0_relu,
                         std::vector<float> conv0_scales, round_mode conv0_round_mode) {
  return conv(src, wei, bia, sz_stride, sz_padding, nullptr, nullptr, dst, conv0_relu, conv0_scales,
              conv0_round_mode);
}

// ------------------------------------------------------------------------------ ext API
namespace ext {

void *device_data(memory &m) { return mirror(m); }
void to_device(memory &m, void *stream) {
  cuda_or_exit(df_h2d(mirror(m), m.data(), m.buffer_size(), stream), "to_device");
}
void to_host(memory &m, void *stream) { cuda_or_exit(df_d2h(m.data(), mirror(m), m.buffer_size(), stream), "to_host"); }
void submit_device(op &o, void *stream) {
  detail::device_op *d = dynamic_cast<detail::device_op *>(&o);
  if (!d) error_and_exit("submit_device: not a B200 op");
  d->launch(stream);
}
void sync(void *stream) { cuda_or_exit(df_stream_sync(stream), "sync"); }
int launches_per_submit(op &o) {
  detail::device_op *d = dynamic_cast<detail::device_op *>(&o);
  return d ? d->launches() : 0;
}
void pin(memory &m) {
  detail::memory_state *st = detail::state_of(m);
  if (!st->pinned && df_host_register(m.data(), m.buffer_size()) == 0) st->pinned = true;
}
std::unique_ptr<op> conv_sharded(const std::vector<int> &devices, const std::unique_ptr<memory> &src,
                                 const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
                                 std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                                 const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1,
                                 std::unique_ptr<memory> &dst, bool conv0_relu, std::vector<float> conv0_scales,
                                 round_mode conv0_round_mode, bool conv1_relu, std::vector<float> conv1_scales,
                                 round_mode conv1_round_mode) {
  return std::unique_ptr<op>(new sharded_conv_op(devices, src, wei, bia, sz_stride, sz_padding, dst, conv0_scales, conv1_scales,
                                                 wei1x1, bia1x1, conv0_relu, conv1_relu, conv0_round_mode, conv1_round_mode));
}
std::unique_ptr<op> concat_conv(const std::vector<std::unique_ptr<memory>> &srcs, bool concat_relu,
                                const std::unique_ptr<memory> &wei, const std::unique_ptr<memory> &bia,
                                std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                                const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1,
                                std::unique_ptr<memory> &dst, bool conv0_relu, std::vector<float> conv0_scales,
                                round_mode conv0_round_mode, bool conv1_relu, std::vector<float> conv1_scales,
                                round_mode conv1_round_mode) {
  // the concat half of the checks (jit_concat_kernel::init_conf, src/jit_concat_kernel.cc:130-197, u8 only here)
  if (srcs.empty() || !dst) error_and_exit("Init Concat op failed!");
  const memory::dims d0 = srcs[0]->actual_dims();
  int channels = 0;
  for (auto &m : srcs) {
    const memory::dims d = m->actual_dims();
    if (m->dim_format() != memory::format::nhwc || m->data_type() != memory::dtype::u8 || d.size() != 4 || d[0] != d0[0] ||
        d[1] != d0[1] || d[2] != d0[2] || d[3] % 16) {
      info("concat_conv inputs must be nhwc u8 with equal N, H, W and channels in multiples of 16");
      error_and_exit("Init Concat op failed!");
    }
    channels += d[3];
  }
  std::unique_ptr<memory> shape(new memory(memory::nchw_dims{d0[0], channels, d0[1], d0[2]}, memory::format::nhwc, memory::dtype::u8));
  return std::unique_ptr<op>(new concat_conv_op(shape, srcs, concat_relu, wei, bia, sz_stride, sz_padding, dst, conv0_scales,
                                                conv1_scales, wei1x1, bia1x1, conv0_relu, conv1_relu, conv0_round_mode,
                                                conv1_round_mode));
}
std::unique_ptr<op> conv_pool(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                              const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                              std::unique_ptr<memory> &conv_dst, std::unique_ptr<memory> &pool_dst, pool_kind kind,
                              std::array<int, 2> pool_kernel, std::array<int, 2> pool_stride, std::array<int, 2> pool_padding,
                              bool conv_relu, std::vector<float> conv_scales, round_mode conv_round_mode,
                              round_mode pool_round_mode) {
  return std::unique_ptr<op>(new conv_pool_op(src, wei, bia, sz_stride, sz_padding, conv_dst, pool_dst, (int)kind, pool_kernel,
                                              pool_stride, pool_padding, conv_relu, conv_scales, conv_round_mode, pool_round_mode));
}
std::unique_ptr<op> pool(const std::unique_ptr<memory> &src, std::unique_ptr<memory> &dst, pool_kind kind,
                         std::array<int, 2> pool_kernel, std::array<int, 2> pool_stride, std::array<int, 2> pool_padding,
                         round_mode pool_round_mode) {
  return std::unique_ptr<op>(new pool_op(src, dst, (int)kind, pool_kernel, pool_stride, pool_padding, pool_round_mode));
}
std::unique_ptr<op> conv_sum(const std::unique_ptr<memory> &src, const std::unique_ptr<memory> &wei,
                             const std::unique_ptr<memory> &bia, std::array<int, 2> sz_stride, std::array<int, 2> sz_padding,
                             const std::unique_ptr<memory> &wei1x1, const std::unique_ptr<memory> &bia1x1,
                             const std::unique_ptr<memory> &residual, std::unique_ptr<memory> &dst, bool conv0_relu,
                             std::vector<float> conv0_scales, round_mode conv0_round_mode, bool conv1_relu,
                             std::vector<float> conv1_scales, round_mode conv1_round_mode) {
  return std::unique_ptr<op>(new conv_sum_op(src, wei, bia, sz_stride, sz_padding, wei1x1, bia1x1, residual, dst, conv0_relu,
                                             conv0_scales, conv0_round_mode, conv1_relu, conv1_scales, conv1_round_mode));
}
bool concat_conv_is_fused(op &o) {
  concat_conv_op *c = dynamic_cast<concat_conv_op *>(&o);
  return c && c->fused();
}
void sharded_upload(op &o) {
  sharded_conv_op *s = dynamic_cast<sharded_conv_op *>(&o);
  if (!s) error_and_exit("sharded_upload: not a sharded op");
  s->upload();
}
void sharded_download(op &o) {
  sharded_conv_op *s = dynamic_cast<sharded_conv_op *>(&o);
  if (!s) error_and_exit("sharded_download: not a sharded op");
  s->download();
}
void sharded_sync(op &o) {
  sharded_conv_op *s = dynamic_cast<sharded_conv_op *>(&o);
  if (!s) error_and_exit("sharded_sync: not a sharded op");
  s->sync_all();
}
void release(std::unique_ptr<op> &o) {
  if (o) detail::release_resources(o.get());
  o.reset();
}

}  // namespace ext
}  // namespace deepfusion
