// df_runtime.cu -- device / memory / stream / event plumbing of the C-ABI (include/dfcuda.h).
#include <string.h>

#include "df_common.cuh"

namespace df {
char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace df

extern "C" {

const char* df_last_error(void) { return df::last_error_buf(); }
const char* df_version(void) { return "deep-fusion_b200 0.1 (sm_100a)"; }

int df_device_count(int* count) {
  if (!count) return df::fail(DF_E_INVALID, "df_device_count: null");
  DF_CUDA(cudaGetDeviceCount(count));
  return 0;
}
int df_set_device(int device) {
  DF_CUDA(cudaSetDevice(device));
  return 0;
}
int df_get_device(int* device) {
  if (!device) return df::fail(DF_E_INVALID, "df_get_device: null");
  DF_CUDA(cudaGetDevice(device));
  return 0;
}
int df_device_sm_count(int* sms) {
  int dev;
  DF_CUDA(cudaGetDevice(&dev));
  DF_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}
int df_malloc(size_t bytes, void** p) {
  if (!p) return df::fail(DF_E_INVALID, "df_malloc: null");
  DF_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  return 0;
}
int df_free(void* p) {
  DF_CUDA(cudaFree(p));
  return 0;
}
int df_memset(void* p, int v, size_t bytes, void* stream) {
  DF_CUDA(cudaMemsetAsync(p, v, bytes, (cudaStream_t)stream));
  return 0;
}
int df_host_register(void* p, size_t bytes) {
  DF_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return 0;
}
int df_host_unregister(void* p) {
  DF_CUDA(cudaHostUnregister(p));
  return 0;
}
int df_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  DF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
int df_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  DF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
int df_stream_create(void** s) {
  cudaStream_t st;
  DF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = st;
  return 0;
}
int df_stream_sync(void* s) {
  DF_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  return 0;
}
int df_stream_destroy(void* s) {
  DF_CUDA(cudaStreamDestroy((cudaStream_t)s));
  return 0;
}
int df_event_create(void** e) {
  cudaEvent_t ev;
  DF_CUDA(cudaEventCreate(&ev));
  *e = ev;
  return 0;
}
int df_event_record(void* e, void* s) {
  DF_CUDA(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s));
  return 0;
}
int df_event_record_node(void* e, void* s) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  DF_CUDA(cudaStreamIsCapturing((cudaStream_t)s, &cs));
  DF_CUDA(cudaEventRecordWithFlags((cudaEvent_t)e, (cudaStream_t)s, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
  return 0;
}
int df_stream_wait_event(void* s, void* e) {
  DF_CUDA(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)e, 0));
  return 0;
}
int df_event_elapsed_ms(void* a, void* b, float* ms) {
  DF_CUDA(cudaEventSynchronize((cudaEvent_t)b));
  DF_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return 0;
}
int df_event_destroy(void* e) {
  DF_CUDA(cudaEventDestroy((cudaEvent_t)e));
  return 0;
}

// ---- CUDA graphs: replay a captured sequence of df_* calls with one launch (launch-bound loops)
int df_graph_begin(void* s) {
  if (!s) return df::fail(DF_E_INVALID, "graph capture needs an explicit stream (df_stream_create)");
  DF_CUDA(cudaStreamBeginCapture((cudaStream_t)s, cudaStreamCaptureModeThreadLocal));
  return 0;
}
int df_graph_end(void* s, void** graph_exec) {
  if (!graph_exec) return df::fail(DF_E_INVALID, "graph: null out");
  *graph_exec = nullptr;
  cudaGraph_t g = nullptr;
  DF_CUDA(cudaStreamEndCapture((cudaStream_t)s, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return df::fail((int)e, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  *graph_exec = ge;
  return 0;
}
int df_graph_launch(void* graph_exec, void* s) {
  DF_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)s));
  return 0;
}
int df_graph_destroy(void* graph_exec) {
  if (graph_exec) DF_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return 0;
}

}  // extern "C"
