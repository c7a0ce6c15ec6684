// deepfusion_c.cc -- C wrappers over the C++ API (include/deepfusion_c.h).
#include "deepfusion_c.h"

#include <vector>

#include "deepfusion.h"
#include "deepfusion_ext.h"

using namespace deepfusion;

struct dfh_memory {
  std::unique_ptr<memory> m;
};
struct dfh_op {
  std::unique_ptr<op> o;
};

extern "C" {

dfh_memory *dfh_memory_create_nchw(const int d[4], int format, int dtype, int alignment) {
  memory::nchw_dims dm = {d[0], d[1], d[2], d[3]};
  dfh_memory *h = new dfh_memory();
  h->m.reset(new memory(dm, (memory::format)format, (memory::dtype)dtype, alignment > 0 ? alignment : 4096));
  return h;
}
dfh_memory *dfh_memory_create(const int *dims, int ndims, int format, int dtype, int alignment) {
  memory::dims dm(dims, dims + ndims);
  dfh_memory *h = new dfh_memory();
  h->m.reset(new memory(dm, (memory::format)format, (memory::dtype)dtype, alignment > 0 ? alignment : 4096));
  return h;
}
void *dfh_memory_data(dfh_memory *m) { return m->m->data(); }
size_t dfh_memory_bytes(dfh_memory *m) { return m->m->buffer_size(); }
void *dfh_memory_device(dfh_memory *m) { return ext::device_data(*m->m); }
void dfh_memory_pin(dfh_memory *m) { ext::pin(*m->m); }
void dfh_memory_to_device(dfh_memory *m) { ext::to_device(*m->m); }
void dfh_memory_to_host(dfh_memory *m) { ext::to_host(*m->m); }
void dfh_memory_destroy(dfh_memory *m) { delete m; }

dfh_op *dfh_concat_create(dfh_memory *const *srcs, int n, dfh_memory *dst, int post_relu) {
  // the C++ factory takes a vector of unique_ptr: lend the pointers for the duration of the call
  std::vector<std::unique_ptr<memory>> v;
  for (int i = 0; i < n; ++i) v.emplace_back(srcs[i]->m.release());
  dfh_op *h = new dfh_op();
  h->o = concat(v, dst->m, post_relu != 0);
  for (int i = 0; i < n; ++i) srcs[i]->m.reset(v[i].release());
  return h;
}

dfh_op *dfh_conv_create(dfh_memory *src, dfh_memory *wei, dfh_memory *bia, const int stride[2], const int padding[2],
                        dfh_memory *wei1x1, dfh_memory *bia1x1, dfh_memory *dst, int conv0_relu,
                        const float *s0, int n0, int r0, int conv1_relu, const float *s1, int n1, int r1) {
  static const std::unique_ptr<memory> none;
  std::vector<float> sc0(s0, s0 + (n0 > 0 ? n0 : 0)), sc1;
  if (sc0.empty()) sc0.push_back(1.f);
  if (s1 && n1 > 0) sc1.assign(s1, s1 + n1);
  if (sc1.empty()) sc1.push_back(1.f);
  dfh_op *h = new dfh_op();
  const std::array<int, 2> st = {stride[0], stride[1]}, pd = {padding[0], padding[1]};
  if (wei1x1)
    h->o = conv(src->m, wei->m, bia ? bia->m : none, st, pd, wei1x1->m, bia1x1 ? bia1x1->m : none, dst->m,
                conv0_relu != 0, sc0, (round_mode)r0, conv1_relu != 0, sc1, (round_mode)r1);
  else
    h->o = conv(src->m, wei->m, bia ? bia->m : none, st, pd, dst->m, conv0_relu != 0, sc0, (round_mode)r0);
  return h;
}

dfh_op *dfh_conv_sharded_create(const int *devices, int n_devices, dfh_memory *src, dfh_memory *wei, dfh_memory *bia,
                                const int stride[2], const int padding[2], dfh_memory *wei1x1, dfh_memory *bia1x1,
                                dfh_memory *dst, int conv0_relu, const float *s0, int n0, int r0, int conv1_relu,
                                const float *s1, int n1, int r1) {
  static const std::unique_ptr<memory> none;
  std::vector<float> sc0(s0, s0 + (n0 > 0 ? n0 : 0)), sc1;
  if (sc0.empty()) sc0.push_back(1.f);
  if (s1 && n1 > 0) sc1.assign(s1, s1 + n1);
  if (sc1.empty()) sc1.push_back(1.f);
  dfh_op *h = new dfh_op();
  const std::array<int, 2> st = {stride[0], stride[1]}, pd = {padding[0], padding[1]};
  h->o = ext::conv_sharded(std::vector<int>(devices, devices + n_devices), src->m, wei->m, bia ? bia->m : none, st, pd,
                           wei1x1 ? wei1x1->m : none, bia1x1 ? bia1x1->m : none, dst->m, conv0_relu != 0, sc0, (round_mode)r0,
                           conv1_relu != 0, sc1, (round_mode)r1);
  return h;
}
dfh_op *dfh_concat_conv_create(dfh_memory *const *srcs, int n, int concat_relu, dfh_memory *wei, dfh_memory *bia,
                               const int stride[2], const int padding[2], dfh_memory *wei1x1, dfh_memory *bia1x1,
                               dfh_memory *dst, int conv0_relu, const float *s0, int n0, int r0, int conv1_relu,
                               const float *s1, int n1, int r1) {
  static const std::unique_ptr<memory> none;
  std::vector<float> sc0(s0, s0 + (n0 > 0 ? n0 : 0)), sc1;
  if (sc0.empty()) sc0.push_back(1.f);
  if (s1 && n1 > 0) sc1.assign(s1, s1 + n1);
  if (sc1.empty()) sc1.push_back(1.f);
  std::vector<std::unique_ptr<memory>> v;  // lend the pointers for the duration of the call (see dfh_concat_create)
  for (int i = 0; i < n; ++i) v.emplace_back(srcs[i]->m.release());
  dfh_op *h = new dfh_op();
  const std::array<int, 2> st = {stride[0], stride[1]}, pd = {padding[0], padding[1]};
  h->o = ext::concat_conv(v, concat_relu != 0, wei->m, bia ? bia->m : none, st, pd, wei1x1 ? wei1x1->m : none,
                          bia1x1 ? bia1x1->m : none, dst->m, conv0_relu != 0, sc0, (round_mode)r0, conv1_relu != 0, sc1,
                          (round_mode)r1);
  for (int i = 0; i < n; ++i) srcs[i]->m.reset(v[i].release());
  return h;
}
dfh_op *dfh_conv_pool_create(dfh_memory *src, dfh_memory *wei, dfh_memory *bia, const int stride[2], const int padding[2],
                             dfh_memory *conv_dst, dfh_memory *pool_dst, int kind, const int pk[2], const int ps[2],
                             const int pp[2], int conv_relu, const float *s0, int n0, int r0, int pool_round) {
  static const std::unique_ptr<memory> none;
  std::vector<float> sc0(s0, s0 + (n0 > 0 ? n0 : 0));
  if (sc0.empty()) sc0.push_back(1.f);
  dfh_op *h = new dfh_op();
  h->o = ext::conv_pool(src->m, wei->m, bia ? bia->m : none, {stride[0], stride[1]}, {padding[0], padding[1]}, conv_dst->m,
                        pool_dst->m, (ext::pool_kind)kind, {pk[0], pk[1]}, {ps[0], ps[1]}, {pp[0], pp[1]}, conv_relu != 0, sc0,
                        (round_mode)r0, (round_mode)pool_round);
  return h;
}
dfh_op *dfh_pool_create(dfh_memory *src, dfh_memory *dst, int kind, const int pk[2], const int ps[2], const int pp[2], int pool_round) {
  dfh_op *h = new dfh_op();
  h->o = ext::pool(src->m, dst->m, (ext::pool_kind)kind, {pk[0], pk[1]}, {ps[0], ps[1]}, {pp[0], pp[1]}, (round_mode)pool_round);
  return h;
}
dfh_op *dfh_conv_sum_create(dfh_memory *src, dfh_memory *wei, dfh_memory *bia, const int stride[2], const int padding[2],
                            dfh_memory *wei1x1, dfh_memory *bia1x1, dfh_memory *residual, dfh_memory *dst, int conv0_relu,
                            const float *s0, int n0, int r0, int conv1_relu, const float *s1, int n1, int r1) {
  static const std::unique_ptr<memory> none;
  std::vector<float> sc0(s0, s0 + (n0 > 0 ? n0 : 0)), sc1;
  if (sc0.empty()) sc0.push_back(1.f);
  if (s1 && n1 > 0) sc1.assign(s1, s1 + n1);
  if (sc1.empty()) sc1.push_back(1.f);
  dfh_op *h = new dfh_op();
  h->o = ext::conv_sum(src->m, wei->m, bia ? bia->m : none, {stride[0], stride[1]}, {padding[0], padding[1]},
                       wei1x1 ? wei1x1->m : none, bia1x1 ? bia1x1->m : none, residual->m, dst->m, conv0_relu != 0, sc0,
                       (round_mode)r0, conv1_relu != 0, sc1, (round_mode)r1);
  return h;
}
int dfh_concat_conv_is_fused(dfh_op *op) { return ext::concat_conv_is_fused(*op->o) ? 1 : 0; }
void dfh_sharded_upload(dfh_op *op) { ext::sharded_upload(*op->o); }
void dfh_sharded_sync(dfh_op *op) { ext::sharded_sync(*op->o); }
void dfh_sharded_download(dfh_op *op) { ext::sharded_download(*op->o); }

void dfh_op_submit(dfh_op *op) { op->o->submit(); }
void dfh_op_submit_device(dfh_op *op, void *stream) { ext::submit_device(*op->o, stream); }
int dfh_op_launches(dfh_op *op) { return ext::launches_per_submit(*op->o); }
void dfh_sync(void *stream) { ext::sync(stream); }
void dfh_op_destroy(dfh_op *op) {
  if (op) ext::release(op->o);
  delete op;
}

}  // extern "C"
