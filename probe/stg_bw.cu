// stg_bw.cu -- how fast can one SM push global stores (L2-resident footprint) for the patterns the conv
// epilogue can produce?  16 warps per SM store continuously; reports bytes/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

// P: 0 = STG.128, lane l -> row l, 16 B (32 lines / instr)           [old epilogue]
//    1 = STG.128, 8 rows x 64 B                                       [row-pair fragments, u8]
//    2 = STG.128, 4 rows x 128 B
//    3 = STG.128, 512 B contiguous
//    4 = STG.256, 8 rows x 128 B
//    5 = STG.256, lane l -> row l, 32 B
template <int P>
__global__ void __launch_bounds__(512, 1) k(uint8_t* out, int iters, long long* cyc, int pitch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = out + (size_t)blockIdx.x * (128 * pitch);   // 128 rows per SM
  uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // each warp owns rows (warp%4)*32.. and column block (warp/4)*128 within a 512-byte row
    const int col0 = (warp >> 2) * 128, row0 = (warp & 3) * 32;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      uint8_t* p;
      if (P == 0) p = base + (size_t)(row0 + lane) * pitch + col0 + s * 16;
      else if (P == 1) p = base + (size_t)(row0 + (s & 3) * 8 + (lane >> 2)) * pitch + col0 + (s >> 2) * 64 + (lane & 3) * 16;
      else if (P == 2) p = base + (size_t)(row0 + s * 4 + (lane >> 3)) * pitch + col0 + (lane & 7) * 16;
      else if (P == 3) p = base + (size_t)(row0 + s * 4 + (lane >> 3)) * 128 + warp * 4096 + (lane & 7) * 16;
      else if (P == 4) p = base + (size_t)(row0 + (s & 3) * 8 + (lane >> 2)) * pitch + col0 + (lane & 3) * 32;
      else p = base + (size_t)(row0 + lane) * pitch + col0 + (s & 3) * 32;
      v.x += it;
      if (P >= 4) asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      else *reinterpret_cast<uint4*>(p) = v;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int P> void run(const char* name, int warps) {
  uint8_t* out; long long* cyc;
  const int pitch = 512;
  CK(cudaMalloc(&out, (size_t)148 * 128 * pitch)); CK(cudaMalloc(&cyc, 148 * 8));
  const int iters = 2000;
  k<P><<<148, warps * 32>>>(out, iters, cyc, pitch); CK(cudaDeviceSynchronize());
  k<P><<<148, warps * 32>>>(out, iters, cyc, pitch); CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  const double bytes = (double)iters * 8 * warps * 32 * (P >= 4 ? 32 : 16);
  printf("%-48s warps=%2d: %6.1f B/clk/SM  (%5.1f clk per warp store)\n", name, warps, bytes / c, (double)c / (iters * 8.0 * warps));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 16}) {
    run<0>("STG.128  32 rows x 16 B", w);
    run<1>("STG.128   8 rows x 64 B", w);
    run<2>("STG.128   4 rows x 128 B", w);
    run<3>("STG.128   512 B contiguous", w);
    run<4>("STG.256   8 rows x 128 B", w);
    run<5>("STG.256  32 rows x 32 B", w);
  }
  return 0;
}
