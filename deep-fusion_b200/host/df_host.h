// df_host.h -- private declarations of the C++ host layer (not installed).
#pragma once
#include <stdio.h>

#include "deepfusion.h"
#include "deepfusion_ext.h"
#include "dfcuda.h"

namespace deepfusion {

// log / check macros with the reference's wording and exit behaviour (util/log.h:26-65)
#define DF_FILENAME (__builtin_strrchr(__FILE__, '/') ? __builtin_strrchr(__FILE__, '/') + 1 : __FILE__)
#define DF_LOG(stream, type, fmt, ...) \
  fprintf(stream, "[" #type " %s %s:%d] >> " fmt "\n", __TIME__, DF_FILENAME, __LINE__, ##__VA_ARGS__)
#define info(fmt, ...) DF_LOG(stdout, INFO, fmt, ##__VA_ARGS__)
#define warning(fmt, ...) DF_LOG(stdout, WARNING, fmt, ##__VA_ARGS__)
#define error_and_exit(fmt, ...)               \
  {                                            \
    DF_LOG(stderr, ERROR, fmt, ##__VA_ARGS__); \
    exit(EXIT_FAILURE);                        \
  }
#define check_eq(a, b) \
  if (!((a) == (b))) error_and_exit("Check " #a " == " #b " Failed!")

namespace detail {

// device mirror of one memory (side table keyed by the memory's address: the public header is the
// reference's, unchanged, and has no room for it)
struct memory_state {
  void *dev = nullptr;
  bool pinned = false;
};
memory_state *state_of(memory &m);

// what an op owns; released by ext::release(), address reuse or process exit (op has no virtual destructor)
struct op_resources {
  virtual ~op_resources() {}
};
void adopt_resources(const op *o, op_resources *r);
bool release_resources(const op *o);

// every B200 op: launch on device mirrors, asynchronously
class device_op : public op {
public:
  virtual void launch(void *stream) = 0;
  virtual int launches() const = 0;
};

size_t dtype_size(memory::dtype dt);           // util/memory.cc:42-56
int conv_output_size(int image, int kernel, int stride, int padding);  // util/math_func.cc:22-24
bool profiling_enabled();                      // env DEEPFUSION_VERBOSE / DEEPFUSION_PROFILE

}  // namespace detail
}  // namespace deepfusion
