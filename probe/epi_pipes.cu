// epi_pipes.cu -- issue throughput of the epilogue's instruction candidates on sm_100a.
// Every thread runs UNROLL independent chains for ITERS iterations; 4..16 warps per SM (one CTA
// per SM).  Reports warp-instructions per clock per SM (4 SMSPs -> 4.0 is full rate).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

constexpr int UNROLL = 16, ITERS = 512;

template <int OP>
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t r[UNROLL];
  float f[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) { r[i] = seed + threadIdx.x * 17 + i * 1000; f[i] = (float)(seed + i) * 0.37f + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (OP == 0) { f[i] = __int2float_rn((int)r[i]); r[i] = __float_as_uint(f[i]) + it; }          // I2F (+IADD)
      if (OP == 1) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.25f)); }             // FADD
      if (OP == 2) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.0001f)); }           // FMUL
      if (OP == 6) { r[i] = (uint32_t)__float2int_rn(f[i]) ; f[i] = __uint_as_float(r[i] | 0x3f000000u); }  // F2I (+LOP)
      if (OP == 7) { r[i] = r[i] + 0x4B400000u; f[i] = __uint_as_float(r[i]) - 12582912.0f; r[i] = __float_as_uint(f[i]) & 0xffff; } // magic I2F
      if (OP == 8) { r[i] = r[i] * 3 + 1; }                                                           // IMAD baseline
    }
    if (OP == 3) {  // FADD2
#pragma unroll
      for (int i = 0; i < UNROLL; i += 2) {
        unsigned long long a, b;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f[i]), "f"(f[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(1.25f), "f"(0.75f));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(f[i + 1]) : "l"(a));
      }
    }
    if (OP == 4) {  // FMUL2
#pragma unroll
      for (int i = 0; i < UNROLL; i += 2) {
        unsigned long long a, b;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f[i]), "f"(f[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(1.0001f), "f"(0.9999f));
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(f[i + 1]) : "l"(a));
      }
    }
    if (OP == 5) {  // cvt.rni + cvt.pack (F2IP.U8.F32): 2 floats -> 2 bytes
#pragma unroll
      for (int i = 0; i < UNROLL; i += 2) {
        int q0 = __float2int_rn(f[i]), q1 = __float2int_rn(f[i + 1]);
        uint32_t pk;
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(pk) : "r"(q1), "r"(q0), "r"(r[i]));
        r[i] = pk; f[i] = __uint_as_float((pk & 0x7fffff) | 0x42000000u); f[i + 1] = f[i] + 1.0f;
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < UNROLL; ++i) acc ^= r[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double instr_per_iter_per_thread, int warps) {
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  k<OP><<<148, warps * 32>>>(out, cyc, 12345u);
  CK(cudaDeviceSynchronize());
  k<OP><<<148, warps * 32>>>(out, cyc, 12345u);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  double winstr = instr_per_iter_per_thread * ITERS * warps;
  printf("%-44s warps=%2d: %7.3f warp-instr/clk/SM (counting %g instr per chain step)\n", name, warps, winstr / (double)c, instr_per_iter_per_thread / UNROLL);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<8>("IMAD (baseline full-rate int)", UNROLL, w);
    run<1>("FADD", UNROLL, w);
    run<2>("FMUL", UNROLL, w);
    run<3>("FADD2 (add.rn.f32x2, per packed instr)", UNROLL / 2, w);
    run<4>("FMUL2 (mul.rn.f32x2, per packed instr)", UNROLL / 2, w);
    run<0>("I2F.S32->F32 (+1 IADD per step)", UNROLL * 2, w);
    run<6>("F2I.RN (+1 LOP per step)", UNROLL * 2, w);
    run<5>("F2IP.U8.F32 pair (+~3 ALU per pair)", UNROLL / 2 * 4, w);
    run<7>("magic I2F: IADD + FADD (+1 LOP)", UNROLL * 3, w);
  }
  return 0;
}
