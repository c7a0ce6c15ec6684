// tma_stream_bw.cu -- how fast can ONE SM stream weight blocks out of L2 with TMA, as a function of
// box size and number of boxes in flight?  All 148 CTAs stream the same 144 KB region (L2 hits after
// the first touch), like the W0 ring of the fused conv.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../deep-fusion_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void wait_mode(uint32_t bar, uint32_t parity, int mode) {
  if (mode == 0) { mbar_wait(bar, parity); return; }                       // try_wait loop (may suspend)
  if (mode == 1) { while (!mbar_test_wait(bar, parity)) {} return; }       // test_wait spin
  // try_wait with a short suspend-time hint
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(20u) : "memory");
  }
}
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int box_rows, int stages, int n_boxes_total, int region_rows, long long* cyc, int mode, int group) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&full[i]), 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t box_bytes = box_rows * 128;  // one "stage" = `group` boxes on one barrier
    long long t0 = clock64();
    int issued = 0, done = 0;
    const int n_stage_loads = n_boxes_total / group;
    auto issue = [&](int s, int idx) {
      mbar_expect_tx(smem_u32(&full[s]), box_bytes * group);
      for (int b = 0; b < group; ++b)
        tma_load_2d(smem_u32(smem + (s * group + b) * box_bytes), &tm, smem_u32(&full[s]), 0, ((idx * group + b) * box_rows) % region_rows);
    };
    for (; issued < stages && issued < n_stage_loads; ++issued) issue(issued, issued);
    while (done < n_stage_loads) {
      const int s = done % stages;
      wait_mode(smem_u32(&full[s]), (done / stages) & 1, mode);
      ++done;
      if (issued < n_stage_loads) { issue(s, issued); ++issued; }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

// plain (non-tensor) bulk copies: cp.async.bulk.shared::cluster.global, `bytes` per instruction
__global__ void __launch_bounds__(128, 1) kb(const uint8_t* src, int bytes, int stages, int n_total, int region_bytes, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&full[i]), 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int issued = 0, done = 0;
    auto issue = [&](int s, int idx) {
      mbar_expect_tx(smem_u32(&full[s]), bytes);
      bulk_load(smem_u32(smem + s * bytes), src + ((size_t)idx * bytes) % region_bytes, bytes, smem_u32(&full[s]));
    };
    for (; issued < stages && issued < n_total; ++issued) issue(issued, issued);
    while (done < n_total) {
      const int s = done % stages;
      while (!mbar_test_wait(smem_u32(&full[s]), (done / stages) & 1)) {}
      ++done;
      if (issued < n_total) { issue(s, issued); ++issued; }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

// burst: issue `n` copies (each on its own barrier) back to back, then wait for all of them
__global__ void __launch_bounds__(128, 1) kburst(const uint8_t* src, int bytes, int n, int reps, long long* cyc, long long* cyc_issue) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&full[i]), 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0, tot_issue = 0;
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        mbar_expect_tx(smem_u32(&full[i]), bytes);
        bulk_load(smem_u32(smem + i * bytes), src + (size_t)i * bytes, bytes, smem_u32(&full[i]));
      }
      long long t1 = clock64();
      for (int i = 0; i < n; ++i) while (!mbar_test_wait(smem_u32(&full[i]), r & 1)) {}
      long long t2 = clock64();
      tot += t2 - t0; tot_issue += t1 - t0;
    }
    cyc[blockIdx.x] = tot / reps; cyc_issue[blockIdx.x] = tot_issue / reps;
  }
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int region_rows = 1152;  // 144 KB of 128-byte rows
  uint8_t* d; CK(cudaMalloc(&d, region_rows * 128)); CK(cudaMemset(d, 1, region_rows * 128));
  long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int bytes : {4096, 16384, 32768, 49152, 65536}) {
    for (int stages : {1, 2, 3}) {
      if ((long)stages * bytes > 190 * 1024) continue;
      const int n_total = 4 * 1024 * 1024 / bytes;
      const int region = 147456 / bytes * bytes;
      kb<<<148, 128, 200 * 1024>>>(d, bytes, stages, n_total, region ? region : bytes, cyc); CK(cudaDeviceSynchronize());
      kb<<<148, 128, 200 * 1024>>>(d, bytes, stages, n_total, region ? region : bytes, cyc); CK(cudaDeviceSynchronize());
      std::vector<long long> c(148); CK(cudaMemcpy(c.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
      long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
      printf("bulk copy %6d B x %d in flight: %6.1f B/clk/SM  (%.0f cycles per copy)\n", bytes, stages, (double)n_total * bytes / mx, (double)mx / n_total);
    }
  }
  CK(cudaFuncSetAttribute(kburst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* cyc2; CK(cudaMalloc(&cyc2, 148 * 8));
  for (int bytes : {4096, 16384}) for (int n : {1, 2, 4, 8}) {
    if ((long)n * bytes > 144 * 1024) continue;
    kburst<<<148, 128, 200 * 1024>>>(d, bytes, n, 200, cyc, cyc2); CK(cudaDeviceSynchronize());
    long long a, b; CK(cudaMemcpy(&a, cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&b, cyc2, 8, cudaMemcpyDeviceToHost));
    printf("burst of %d x %5d B: issue %5lld cycles, all landed after %5lld cycles  (%.1f B/clk/SM)\n", n, bytes, b, a, (double)n * bytes / a);
  }
  for (int box_rows : {128}) {
    CUtensorMap tm;
    cuuint64_t gd[2] = {128, (cuuint64_t)region_rows}; cuuint64_t gs[1] = {128};
    cuuint32_t box[2] = {128, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    for (int stages : {2}) {
      const int box_bytes = box_rows * 128;
      if ((long)stages * box_bytes > 190 * 1024) continue;
      const int n_boxes = 4 * 1024 * 1024 / box_bytes;  // 4 MB per CTA
      for (int group : {1, 2, 4}) {
        const int mode = 1, grid = 148;
        if ((long)stages * group * box_bytes > 190 * 1024) continue;
        k<<<grid, 128, 200 * 1024>>>(tm, box_rows, stages, n_boxes, region_rows, cyc, mode, group); CK(cudaDeviceSynchronize());
        k<<<grid, 128, 200 * 1024>>>(tm, box_rows, stages, n_boxes, region_rows, cyc, mode, group); CK(cudaDeviceSynchronize());
        std::vector<long long> c(grid); CK(cudaMemcpy(c.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
        double bpc = (double)n_boxes * box_bytes / (double)mx;
        printf("box %3d rows (%5d B), %d boxes per barrier, %d barriers in flight: %6.1f B/clk/SM  (%.0f cycles per box, %.0f per barrier)\n", box_rows, box_bytes, group, stages, bpc, (double)mx / n_boxes, (double)mx / n_boxes * group);
      }
    }
  }
  return 0;
}
