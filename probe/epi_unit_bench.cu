// epi_unit_bench.cu -- the conv1 epilogue unit (128 rows x 128 columns: TMEM s32 -> scale -> u8 ->
// global) in isolation: 16 epilogue warps in two groups, as in conv_fused_kernel, no MMA running.
// Variants isolate the cost of each ingredient.  Reports cycles per unit per group and the
// resulting elements/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../deep-fusion_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

struct Consts { float scale[512]; };

__device__ __forceinline__ uint32_t pack_u8x4(const float* t) {
  int q0 = __float2int_rn(t[0]), q1 = __float2int_rn(t[1]), q2 = __float2int_rn(t[2]), q3 = __float2int_rn(t[3]);
  uint32_t hi, lo;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(q3), "r"(q2));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(q1), "r"(q0), "r"(hi));
  return lo;
}
__device__ __forceinline__ void fast4_packed(const uint32_t* acc, int k, const float4 c, const float4 s, float* t) {
  const float f0 = __int_as_float((int)acc[0] + k), f1 = __int_as_float((int)acc[1] + k);
  const float f2 = __int_as_float((int)acc[2] + k), f3 = __int_as_float((int)acc[3] + k);
  unsigned long long a, b, cc, dd, s0, s1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f0), "f"(f1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(f2), "f"(f3));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(c.z), "f"(c.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(s0) : "f"(s.x), "f"(s.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(s1) : "f"(s.z), "f"(s.w));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(cc));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(b), "l"(dd));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(s0));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(b) : "l"(b), "l"(s1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t[0]), "=f"(t[1]) : "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t[2]), "=f"(t[3]) : "l"(b));
}
__device__ __forceinline__ void fast4_scalar(const uint32_t* acc, int k, const float4 c, const float4 s, float* t) {
  t[0] = __fmul_rn(__fadd_rn(__int_as_float((int)acc[0] + k), c.x), s.x);
  t[1] = __fmul_rn(__fadd_rn(__int_as_float((int)acc[1] + k), c.y), s.y);
  t[2] = __fmul_rn(__fadd_rn(__int_as_float((int)acc[2] + k), c.z), s.z);
  t[3] = __fmul_rn(__fadd_rn(__int_as_float((int)acc[3] + k), c.w), s.w);
}

// variant bits: 1 = C from LDS (else register constant), 2 = scale from constant bank (else register),
//               4 = global stores, 8 = scalar math instead of packed, 16 = scale from LDS too
template <int V, int S>
__global__ void __launch_bounds__(640, 1) k(const __grid_constant__ Consts cst, int units, uint8_t* out, long long* cyc, int kuni) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sC[512];
  __shared__ __align__(16) float sS[512];
  __shared__ __align__(128) uint8_t stage[16][32 * 64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) { sC[i] = 0.25f * i; sS[i] = cst.scale[i]; }
  if (warp == 3) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0;
  if (warp >= 4) {
    const int ew = warp - 4, group = ew / 8, quarter = warp & 3, half = (ew % 8) >> 2;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int m = quarter * 32 + lane;
    t0 = clock64();
    for (int u = group; u < units; u += 2) {
      const int j = u & 3;
      uint8_t* out_row = out + ((size_t)(blockIdx.x * 128 + m)) * 512;
      const uint32_t t_base = lane_addr + 256 + (u & 1) * 128;
      uint8_t* st = stage[ew];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int gr = i * 2 + half;
        uint32_t acc[32];
        tmem_ld_x32(t_base + gr * 32, acc);
        tmem_ld_wait();
        uint32_t keep[8];
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int col = j * 128 + gr * 32 + sub * 16;
          uint32_t packed[4];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            float4 c, s;
            if (V & 1) c = *reinterpret_cast<const float4*>(&sC[col + g4 * 4]);
            else c = make_float4(1.f, 2.f, 3.f, 4.f);
            if (V & 16) s = *reinterpret_cast<const float4*>(&sS[col + g4 * 4]);
            else if (V & 2) { const int c0 = col + g4 * 4; s = make_float4(cst.scale[c0], cst.scale[c0 + 1], cst.scale[c0 + 2], cst.scale[c0 + 3]); }
            else s = make_float4(0.5f, 0.25f, 0.125f, 0.75f);
            float t[4];
            if (V & 8) fast4_scalar(acc + sub * 16 + g4 * 4, kuni, c, s, t);
            else fast4_packed(acc + sub * 16 + g4 * 4, kuni, c, s, t);
            packed[g4] = pack_u8x4(t);
          }
          if (S == 1) *reinterpret_cast<uint4*>(out_row + col) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          else if (S == 2) { keep[sub * 4 + 0] = packed[0]; keep[sub * 4 + 1] = packed[1]; keep[sub * 4 + 2] = packed[2]; keep[sub * 4 + 3] = packed[3]; }
          else if (S == 3) {
            const int piece = i * 2 + sub;  // 0..3: which 16 B piece of this warp's 64-byte row segment
            *reinterpret_cast<uint4*>(st + lane * 64 + ((piece ^ ((lane >> 1) & 3)) * 16)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
          else if (packed[0] == 0x12345678u && packed[3] == 0x9abcdef0u) out_row[col] = 1;
        }
        if (S == 2) {
          const int col = j * 128 + gr * 32;
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(out_row + col), "r"(keep[0]), "r"(keep[1]), "r"(keep[2]), "r"(keep[3]), "r"(keep[4]), "r"(keep[5]), "r"(keep[6]), "r"(keep[7]) : "memory");
        }
      }
      if (S == 3) {
        // this warp's 32 rows x 64 B (column groups half, half+2 -> two 32 B runs per row) go out as
        // 8 rows per instruction: lane l -> row 8k + l/4, piece l%4
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int r = kk * 8 + (lane >> 2), piece = lane & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(st + r * 64 + ((piece ^ ((r >> 1) & 3)) * 16));
          const int gr = (piece >> 1) * 2 + half;
          uint8_t* dst = out + ((size_t)(blockIdx.x * 128 + quarter * 32 + r)) * 512 + j * 128 + gr * 32 + (piece & 1) * 16;
          *reinterpret_cast<uint4*>(dst) = v;
        }
        __syncwarp();
      }
    }
    if (threadIdx.x == 128) cyc[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 3) tmem_dealloc<512>(tmem);
}

template <int V, int S>
void run(const char* name) {
  static Consts h; for (int i = 0; i < 512; ++i) h.scale[i] = 1.0f / (512 + i);
  uint8_t* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)148 * 128 * 512)); CK(cudaMalloc(&cyc, 148 * 8));
  const int units = 400;
  k<V, S><<<148, 640>>>(h, units, out, cyc, 0x4B000000); CK(cudaDeviceSynchronize());
  k<V, S><<<148, 640>>>(h, units, out, cyc, 0x4B000000); CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  // group 0 processed units/2 units in c cycles while group 1 did the same concurrently
  printf("%-64s: %6.0f cycles per unit per group, %5.1f elements/clk/SM\n", name, (double)c / (units / 2), (double)units * 16384 / c);
  cudaFree(out); cudaFree(cyc);
}

// ---- row-pair fragments: tcgen05.ld.16x256b.x8 gives thread t rows (t/4)+{0,8} and columns 8k+2(t%4)+{0,1},
// k=0..7, of a 16-lane x 64-column block.  With the weight rows permuted so that those 16 columns are 16
// consecutive channels, (a) one set of 16+16 constants serves four rows, (b) each thread owns a 16-byte
// piece of a row and four lanes a 64-byte run, so STG.128 goes out as 8 rows x 64 B without staging.
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// CONSTS: 0 = registers, 1 = C and scale from shared memory (LDS.128, four distinct addresses per warp)
template <int CONSTS, int STORE>
__global__ void __launch_bounds__(640, 1) k2(int units, uint8_t* out, long long* cyc, int kuni) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sC[512];
  __shared__ __align__(16) float sS[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) { sC[i] = 0.25f * i; sS[i] = 1.0f / (512 + i); }
  if (warp == 3) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0;
  if (warp >= 4) {
    const int ew = warp - 4, group = ew / 8, quarter = warp & 3, half = (ew % 8) >> 2;
    const int m4 = lane & 3, r8 = lane >> 2;
    t0 = clock64();
    for (int u = group; u < units; u += 2) {
      const int j = u & 3;
      const int ch0 = j * 128 + half * 64 + m4 * 16;  // this thread's 16 channels
      float4 c[4], s[4];
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        if (CONSTS) { c[g4] = *reinterpret_cast<const float4*>(&sC[ch0 + g4 * 4]); s[g4] = *reinterpret_cast<const float4*>(&sS[ch0 + g4 * 4]); }
        else { c[g4] = make_float4(1.f, 2.f, 3.f, 4.f); s[g4] = make_float4(0.5f, 0.25f, 0.125f, 0.75f); }
      }
#pragma unroll
      for (int h16 = 0; h16 < 2; ++h16) {
        uint32_t acc[32];
        tmem_ld_16x256b_x8(tmem + ((uint32_t)(quarter * 32 + h16 * 16) << 16) + 256 + (u & 1) * 128 + half * 64, acc);
        tmem_ld_wait();
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
          uint32_t packed[4];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            // channels 4*g4 .. 4*g4+3 of the piece = (k = 2*g4, e = 0,1), (k = 2*g4+1, e = 0,1)
            uint32_t a4[4] = {acc[4 * (2 * g4) + 2 * hl], acc[4 * (2 * g4) + 2 * hl + 1], acc[4 * (2 * g4 + 1) + 2 * hl], acc[4 * (2 * g4 + 1) + 2 * hl + 1]};
            float t[4];
            fast4_packed(a4, kuni, c[g4], s[g4], t);
            packed[g4] = pack_u8x4(t);
          }
          const int row = blockIdx.x * 128 + quarter * 32 + h16 * 16 + hl * 8 + r8;
          if (STORE) *reinterpret_cast<uint4*>(out + (size_t)row * 512 + ch0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          else if (packed[0] == 0x12345678u && packed[3] == 0x9abcdef0u) out[row] = 1;
        }
      }
    }
    if (threadIdx.x == 128) cyc[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 3) tmem_dealloc<512>(tmem);
}
template <int CONSTS, int STORE>
void run2(const char* name) {
  uint8_t* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)148 * 128 * 512)); CK(cudaMalloc(&cyc, 148 * 8));
  const int units = 400;
  k2<CONSTS, STORE><<<148, 640>>>(units, out, cyc, 0x4B000000); CK(cudaDeviceSynchronize());
  k2<CONSTS, STORE><<<148, 640>>>(units, out, cyc, 0x4B000000); CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  printf("%-64s: %6.0f cycles per unit per group, %5.1f elements/clk/SM\n", name, (double)c / (units / 2), (double)units * 16384 / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 0>("math only (constants in registers, no stores)");
  run<0, 1>("+ STG.128 per 16 columns (current)");
  run<0, 2>("+ STG.256 per 32 columns");
  run<0, 3>("+ per-warp smem transpose, coalesced STG.128 (8 rows x 64 B per instr)");
  run<1 | 2, 1>("full constants (LDS + const bank) + STG.128 (current kernel)");
  run<1 | 2, 2>("full constants + STG.256");
  run<1 | 2, 3>("full constants + smem transpose");
  run2<0, 0>("16x256b fragments: math only");
  run2<0, 1>("16x256b fragments: + direct STG.128 (8 rows x 64 B per instr)");
  run2<1, 0>("16x256b fragments: constants via LDS, no stores");
  run2<1, 1>("16x256b fragments: constants via LDS + direct STG.128");
  return 0;
}
