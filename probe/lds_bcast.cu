// lds_bcast.cu -- cost of feeding per-column epilogue constants: broadcast LDS.128 / LDS.64 / LDS.32
// from shared memory vs. indexed constant-bank loads (kernel parameters) vs. immediate constant operands.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
constexpr int ITERS = 256, N = 512;
struct Consts { float v[N * 2]; };

template <int MODE>
__global__ void k(const __grid_constant__ Consts cst, float* out, long long* cyc, int stride) {
  __shared__ __align__(16) float s[N * 2];
  for (int i = threadIdx.x; i < N * 2; i += blockDim.x) s[i] = cst.v[i];
  __syncthreads();
  float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    const int base = (it * stride) & (N - 64);
#pragma unroll
    for (int c = 0; c < 64; c += 4) {
      if (MODE == 0) {  // LDS.128 broadcast (all lanes same address)
        float4 b = *reinterpret_cast<const float4*>(&s[base + c]);
        a0 += b.x; a1 += b.y; a2 += b.z; a3 += b.w;
      } else if (MODE == 1) {  // 2 x LDS.64
        float2 b = *reinterpret_cast<const float2*>(&s[base + c]), d = *reinterpret_cast<const float2*>(&s[base + c + 2]);
        a0 += b.x; a1 += b.y; a2 += d.x; a3 += d.y;
      } else if (MODE == 2) {  // 4 x LDS.32
        a0 += s[base + c]; a1 += s[base + c + 1]; a2 += s[base + c + 2]; a3 += s[base + c + 3];
      } else if (MODE == 3) {  // constant bank, run-time (uniform) index
        a0 += cst.v[base + c]; a1 += cst.v[base + c + 1]; a2 += cst.v[base + c + 2]; a3 += cst.v[base + c + 3];
      } else if (MODE == 4) {  // constant bank, compile-time index -> immediate c[][] operands
        a0 += cst.v[c]; a1 += cst.v[c + 1]; a2 += cst.v[c + 2]; a3 += cst.v[c + 3];
      } else if (MODE == 5) {  // no loads: just the 4 FADDs
        a0 += 1.5f; a1 += 2.5f; a2 += 3.5f; a3 += 4.5f;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  static Consts h; for (int i = 0; i < N * 2; ++i) h.v[i] = i * 0.001f;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  k<MODE><<<148, warps * 32>>>(h, out, cyc, 64); CK(cudaDeviceSynchronize());
  k<MODE><<<148, warps * 32>>>(h, out, cyc, 64); CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  double consts = (double)ITERS * 64 * warps;  // per-warp constants consumed
  printf("%-46s warps=%2d: %6.2f constants/clk/SM (per warp; x32 lanes = %7.1f lane-elements/clk/SM)\n", name, warps, consts / c, consts * 32 / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {8, 16}) {
    run<5>("no loads (4 FADD per 4 constants)", w);
    run<0>("LDS.128 broadcast", w);
    run<1>("2 x LDS.64 broadcast", w);
    run<2>("4 x LDS.32 broadcast", w);
    run<3>("constant bank, run-time uniform index", w);
    run<4>("constant bank, compile-time index", w);
  }
  return 0;
}
