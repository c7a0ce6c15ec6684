// deepfusion.h -- public C++ API of deep-fusion, B200 edition.
//
// Source-compatible with the reference's include/deepfusion.h:25-146: the same namespace, type
// names, enumerators, constructors, accessors and factory signatures (including defaults), so code
// written against the reference recompiles against this header and links libdeepfusion.so
// unchanged.  What differs is behind the API: `memory` also owns a lazily created device mirror
// and `op::submit()` runs hand-written sm_100a kernels through the C-ABI in dfcuda.h (host -> device
// copy of the sources, kernel, device -> host copy of the destination, synchronous like the
// reference).  Additive, non-reference entry points live in deepfusion_ext.h.
#pragma once

#include <stdint.h>
#include <stdlib.h>

#include <array>
#include <memory>
#include <vector>

namespace deepfusion {

typedef float f32;
typedef int32_t s32;
typedef int8_t s8;
typedef uint8_t u8;

// kept for users of the reference header (include/deepfusion.h:33-40)
#ifndef DISABLE_COPY_AND_ASSIGN
#define DISABLE_COPY_AND_ASSIGN(classname)          \
private:                                            \
  classname(const classname &) = delete;            \
  classname(const classname &&) = delete;           \
  classname &operator=(const classname &) = delete; \
  classname &operator=(const classname &&) = delete
#endif

struct opdesc {
  int tmp;
};

// rounding of the f32 -> integer conversions in the conv epilogues (reference :46-49)
enum round_mode {
  nearest = 0,  // round half to even (vcvtps2dq {rn-sae})
  down,         // toward -inf        (vcvtps2dq {rd-sae})
};

namespace detail {
struct memory_state;  // host-layer private: device mirror bookkeeping
}

struct memory {
public:
  // reference :53-61
  enum format {
    format_undef = 0,
    x,
    nchw,
    oihw = nchw,
    nhwc,
    OIhw4i16o4i,
    gOIhw4i16o4i,
  };
  typedef std::vector<int> dims;
  typedef std::array<int, 2> pair_dims;
  typedef std::array<int, 4> nchw_dims;

  // reference :66-72
  enum dtype {
    undef = 0,
    f32,
    s32,
    s8,
    u8,
  };

  // Logical dims are always given as N,C,H,W (or O,I,H,W); the buffer is laid out per `fmt`.
  explicit memory(const nchw_dims &dm, const format fmt, const dtype dt, int alignment = 4096);
  // Dims given in the physical order of `fmt` (used for format::x biases).
  explicit memory(const dims &dm, const format fmt, const dtype dt, int alignment = 4096);
  ~memory();

  size_t size();         // number of elements
  size_t buffer_size();  // bytes
  dims actual_dims() { return dims_; }
  nchw_dims std_dims() { return std_dims_; }  // nchw or oihw
  dtype data_type() { return dt_; }
  format dim_format() { return fmt_; }
  void *data() { return data_; }  // HOST pointer, valid for the lifetime of the object

  detail::memory_state *state() { return state_; }

private:
  void allocate_buffer(int alignment);
  void *data_;
  dims dims_;
  nchw_dims std_dims_;
  format fmt_;
  dtype dt_;
  detail::memory_state *state_;

  DISABLE_COPY_AND_ASSIGN(memory);
};

// An operator borrows the memories it was created with (raw pointers captured at creation, as in
// the reference, src/op_conv.h:82-95): they must outlive the op.  submit() is synchronous and
// must not be called concurrently on the same op.
class op {
public:
  explicit op() {}
  virtual ~op() {}
  virtual void submit();

protected:
  virtual void infer() = 0;
  virtual const char *name() = 0;
  DISABLE_COPY_AND_ASSIGN(op);
};

// concat along channels (+ optional ReLU); reference :116-118
std::unique_ptr<op> concat(const std::vector<std::unique_ptr<memory>> &srcs,
                           std::unique_ptr<memory> &dst,
                           bool post_relu = false);

// conv only; reference :121-129
std::unique_ptr<op> conv(const std::unique_ptr<memory> &src,
                         const std::unique_ptr<memory> &wei,
                         const std::unique_ptr<memory> &bia,
                         std::array<int, 2> sz_stride,
                         std::array<int, 2> sz_padding,
                         std::unique_ptr<memory> &dst,
                         bool conv0_relu = false,
                         std::vector<float> conv0_scales = {1.f},
                         round_mode conv0_round_mode = round_mode::nearest);

// conv + ReLU fused with conv1x1 (+ ReLU); reference :132-145
std::unique_ptr<op> conv(const std::unique_ptr<memory> &src,
                         const std::unique_ptr<memory> &wei,
                         const std::unique_ptr<memory> &bia,
                         std::array<int, 2> sz_stride,
                         std::array<int, 2> sz_padding,
                         const std::unique_ptr<memory> &wei1x1,
                         const std::unique_ptr<memory> &bia1x1,
                         std::unique_ptr<memory> &dst,
                         bool conv0_relu = false,
                         std::vector<float> conv0_scales = {1.f},
                         round_mode conv0_round_mode = round_mode::nearest,
                         bool conv1_relu = false,
                         std::vector<float> conv1_scales = {1.f},
                         round_mode conv1_round_mode = round_mode::nearest);
}  // namespace deepfusion
