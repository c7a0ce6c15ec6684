// tmem_ld_bw.cu -- tcgen05.ld (TMEM -> registers) bandwidth per SM: the ceiling of any epilogue.
// W warps (4..16) each loop over tcgen05.ld.32x32b.{x16,x32,x64} of their lane quarter; no math.
#include <cstdio>
#include <cstdlib>
#include "../deep-fusion_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
        "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
        "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
}

template <int X>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cyc, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[X], acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)(((it * 64) + (warp >> 2) * X) & (512 - X));
    if (X == 16) tmem_ld_x16(tmem + col, r);
    if (X == 32) tmem_ld_x32(tmem + col, r);
    if (X == 64) tmem_ld_x64(tmem + col, r);
    tmem_ld_wait();
    acc ^= r[0] ^ r[X / 2] ^ r[X - 1];
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

template <int X>
void run(int warps) {
  long long* cyc; uint32_t* sink;
  CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&sink, 4096));
  const int iters = 4000;
  k<X><<<148, warps * 32>>>(iters, cyc, sink);
  CK(cudaDeviceSynchronize());
  k<X><<<148, warps * 32>>>(iters, cyc, sink);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  double bytes = (double)iters * warps * 32 * X * 4;
  printf("tcgen05.ld.32x32b.x%-2d warps=%2d: %7.1f B/clk/SM = %6.1f s32 elements/clk/SM  (%lld cycles per load per warp)\n", X, warps,
         bytes / (double)c, bytes / 4 / (double)c, c / iters);
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) { run<16>(w); run<32>(w); run<64>(w); }
  return 0;
}
