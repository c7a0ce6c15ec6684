// tma_loop_variants.cu -- why does a sliding-window bulk-copy loop complete only one 16 KB copy per
// ~530-630 cycles while a burst of 8 completes one per ~190?  Variants of the loop structure.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../deep-fusion_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

// variant 0: one thread waits for the oldest copy, then issues one new copy (sliding window)
// variant 1: producer thread issues when a slot is free (empty barrier), consumer thread (other warp) waits
//            full and releases the slot immediately -- the structure of the conv kernel
// variant 2: like 0 but waits for `stages/2` copies and then issues `stages/2` back to back
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, int bytes, int stages, int n_total, int region, int variant, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[16], empty[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); } fence_mbar_init(); }
  __syncthreads();
  auto issue = [&](int s, int idx) {
    mbar_expect_tx(smem_u32(&full[s]), bytes);
    bulk_load(smem_u32(smem + s * bytes), src + ((size_t)idx * bytes) % region, bytes, smem_u32(&full[s]));
  };
  if (variant == 0 || variant == 2) {
    if (threadIdx.x == 0) {
      long long t0 = clock64();
      int issued = 0, done = 0;
      const int batch = variant == 2 ? stages / 2 : 1;
      for (; issued < stages && issued < n_total; ++issued) issue(issued, issued);
      while (done < n_total) {
        for (int b = 0; b < batch && done < n_total; ++b, ++done) {
          const int s = done % stages;
          while (!mbar_test_wait(smem_u32(&full[s]), (done / stages) & 1)) {}
        }
        for (int b = 0; b < batch && issued < n_total; ++b, ++issued) issue(issued % stages, issued);
      }
      cyc[blockIdx.x] = clock64() - t0;
    }
  } else {
    if (threadIdx.x == 0) {  // producer
      long long t0 = clock64();
      for (int i = 0; i < n_total; ++i) {
        const int s = i % stages;
        while (!mbar_test_wait(smem_u32(&empty[s]), ((i / stages) & 1) ^ 1)) {}
        issue(s, i);
      }
      cyc[blockIdx.x] = clock64() - t0;
    } else if (threadIdx.x == 32) {  // consumer
      for (int i = 0; i < n_total; ++i) {
        const int s = i % stages;
        while (!mbar_test_wait(smem_u32(&full[s]), (i / stages) & 1)) {}
        mbar_arrive(smem_u32(&empty[s]));
      }
    }
  }
}

int main() {
  const int region = 147456;
  uint8_t* d; CK(cudaMalloc(&d, region)); CK(cudaMemset(d, 1, region));
  long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const char* names[] = {"sliding window, one thread", "producer + consumer warps", "wait half / issue half"};
  for (int bytes : {16384}) for (int stages : {2, 4, 8}) for (int variant : {0, 1, 2}) for (int grid : {1, 148}) {
    const int n_total = 2 * 1024 * 1024 / bytes;
    k<<<grid, 128, 200 * 1024>>>(d, bytes, stages, n_total, region, variant, cyc); CK(cudaDeviceSynchronize());
    k<<<grid, 128, 200 * 1024>>>(d, bytes, stages, n_total, region, variant, cyc); CK(cudaDeviceSynchronize());
    std::vector<long long> c(grid); CK(cudaMemcpy(c.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
    printf("%6d B x %d stages, %-28s grid %3d: %6.1f B/clk/SM (%4.0f cycles per copy)\n", bytes, stages, names[variant], grid, (double)n_total * bytes / mx, (double)mx / n_total);
  }
  return 0;
}
