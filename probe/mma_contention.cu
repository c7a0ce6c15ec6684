// mma_contention.cu -- what slows down tcgen05.mma issue when other warps of the CTA are busy?
// One warp issues batches of 18 kind::i8 MMAs (M128 N64 K32, all into one accumulator, like
// conv0 of cfg1) and commits; `companions` other warps do one of:
//   0 park at the final barrier         1 spin on mbarrier.try_wait, all 32 lanes
//   2 spin, one lane + __syncwarp       3 FFMA loop         4 LDS.128 broadcast loop
//   5 tcgen05.ld x16 loop on other columns                  6 spin with __nanosleep(64)
// The MMA warp is warp 0 (mma_last=0) or the highest-numbered warp (mma_last=1).
// Prints cycles per batch as seen by the issuing thread (issue + completion).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../deep-fusion_b200/csrc/sm100_ptx.cuh"
using namespace sm100;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

template <int N>
__global__ void __launch_bounds__(416, 1) k(int batches, int mode, int mma_last, long long* out_issue, long long* out_done, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_mma, bar_never;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int mma_warp = mma_last ? nwarps - 1 : 0;
  for (int i = tid; i < (256 + N) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar_mma), 1); mbar_init(smem_u32(&bar_never), 1); fence_mbar_init(); stop = 0; }
  if (warp == mma_warp) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == mma_warp) {
    const uint32_t idesc = make_idesc_i8(128, N, 0, 1);
    const uint64_t da = make_smem_desc(smem_u32(smem), 16, 512, kLayoutSW64);
    const uint64_t db = make_smem_desc(smem_u32(smem + 256 * 128), 16, 512, kLayoutSW64);
    long long t_issue = 0, t_done = 0;
    for (int b = 0; b < batches; ++b) {
      long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < 18; ++i) umma_i8(tmem, da + 2 * (i & 1) + 8 * (i >> 1), db + 2 * (i & 1), idesc, i);
        umma_commit(smem_u32(&bar_mma));
      }
      __syncwarp();
      long long t1 = clock64();
      mbar_wait_warp(smem_u32(&bar_mma), b & 1);
      long long t2 = clock64();
      t_issue += t1 - t0; t_done += t2 - t0;
    }
    if (lane == 0) { out_issue[blockIdx.x] = t_issue / batches; out_done[blockIdx.x] = t_done / batches; stop = 1; }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bar_never));
  } else {
    float acc = tid;
    if (mode == 1) { mbar_wait(smem_u32(&bar_never), 0); }
    else if (mode == 2) { mbar_wait_warp(smem_u32(&bar_never), 0); }
    else if (mode == 6) { mbar_wait_warp<64>(smem_u32(&bar_never), 0); }
    else if (mode == 3) { while (!stop) { for (int i = 0; i < 64; ++i) acc = acc * 1.0001f + 0.5f; } }
    else if (mode == 4) { const float4* p = reinterpret_cast<const float4*>(smem); while (!stop) { for (int i = 0; i < 16; ++i) { float4 v = p[(i * 7) & 63]; acc += v.x + v.y + v.z + v.w; } } }
    else if (mode == 5) { uint32_t r[16]; while (!stop) { tmem_ld_x16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp & 7) * 16, r); tmem_ld_wait(); acc += __uint_as_float(r[3]); } }
    if (acc == 12345.678f) sink[tid] = acc;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == mma_warp) tmem_dealloc<512>(tmem);
}

// Variant kernel: the conv kernel's GEMM1 structure -- 9 taps x (elect { nks MMAs } ; __syncwarp),
// runtime trip counts (no unrolling), optional distinct B block per tap, optional per-tap trace store.
struct V { int nks, wp, swb, b_stride, trace, taps, mode; };
// Third family: start from the FAST loop (18 unrolled MMAs, constant increments) and add one
// runtime ingredient at a time.
struct W { int variant, one, nmma, step_a, step_b; };
__global__ void __launch_bounds__(128, 1) kw_(int batches, W w, long long* out_done) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar_mma), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_i8(128, 64, 0, 1);
    const uint64_t da = make_smem_desc(smem_u32(smem), 16, 512, kLayoutSW64);
    const uint64_t db = make_smem_desc(smem_u32(smem + 64 * 1024), 16, 512, kLayoutSW64);
    long long t_done = 0;
    for (int b = 0; b < batches; ++b) {
      long long t0 = clock64();
      if (elect_one()) {
        if (w.variant == 0) {          // baseline: everything compile-time
#pragma unroll
          for (int i = 0; i < 18; ++i) umma_i8(tmem, da + 2 * (i & 1) + 232 * (i >> 1), db + 2 * (i & 1) + 256 * (i >> 1), idesc, i);
        } else if (w.variant == 1) {   // runtime accumulate flag
#pragma unroll
          for (int i = 0; i < 18; ++i) umma_i8(tmem, da + 2 * (i & 1) + 232 * (i >> 1), db + 2 * (i & 1) + 256 * (i >> 1), idesc, i * w.one);
        } else if (w.variant == 2) {   // runtime trip count, no unrolling
#pragma unroll 1
          for (int i = 0; i < w.nmma; ++i) umma_i8(tmem, da + 2 * (i & 1) + 232 * (i >> 1), db + 2 * (i & 1) + 256 * (i >> 1), idesc, i);
        } else if (w.variant == 3) {   // unrolled, runtime per-tap steps (uniform kernel params)
#pragma unroll
          for (int i = 0; i < 18; ++i) umma_i8(tmem, da + 2 * (i & 1) + w.step_a * (i >> 1), db + 2 * (i & 1) + w.step_b * (i >> 1), idesc, i);
        } else if (w.variant == 4) {   // unrolled, descriptors advanced by accumulation (d += step)
          uint64_t a = da, bq = db;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            umma_i8(tmem, a, bq, idesc, t);
            umma_i8(tmem, a + 2, bq + 2, idesc, 1);
            a += w.step_a; bq += w.step_b;
          }
        } else if (w.variant == 5) {   // runtime 3x3 loops, accumulation steps, inner ks unrolled by 2
          uint64_t arow = da, bq = db;
          uint32_t acc = 0;
#pragma unroll 1
          for (int kh = 0; kh < 3 * w.one; ++kh) {
            uint64_t a = arow;
#pragma unroll 1
            for (int kw = 0; kw < 3 * w.one; ++kw) {
              umma_i8(tmem, a, bq, idesc, acc);
              umma_i8(tmem, a + 2, bq + 2, idesc, 1);
              acc = 1; a += 4; bq += w.step_b;
            }
            arow += w.step_a;
          }
        }
        umma_commit(smem_u32(&bar_mma));
      }
      __syncwarp();
      mbar_wait_warp(smem_u32(&bar_mma), b & 1);
      t_done += clock64() - t0;
    }
    if (lane == 0) out_done[blockIdx.x] = t_done / batches;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

static void run_w() {
  long long* d_done;
  CK(cudaMalloc(&d_done, 148 * 8));
  size_t smem = 170 * 1024;
  CK(cudaFuncSetAttribute(kw_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[] = {"baseline all compile-time", "runtime accumulate flag", "runtime trip count (no unroll)",
                         "unrolled, runtime steps (mul)", "unrolled, steps by accumulation", "runtime 3x3 loops, accumulation steps"};
  for (int v = 0; v < 6; ++v) {
    W w{v, 1, 18, 232, 256};
    kw_<<<148, 128, smem>>>(200, w, d_done);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("W launch error %s\n", cudaGetErrorString(e)); return; }
    long long d; CK(cudaMemcpy(&d, d_done, 8, cudaMemcpyDeviceToHost));
    printf("ingredient %-40s: %6lld cycles per 18 MMAs (ideal 864)\n", names[v], d);
  }
}

__global__ void __launch_bounds__(128, 1) kv(int batches, V v, long long* out_done, unsigned long long* tr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar_mma), 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_i8(128, 64, 0, 1);
    const uint64_t hi = make_smem_desc(0, 16, 8 * v.swb, v.swb == 128 ? kLayoutSW128 : kLayoutSW64);
    const uint32_t a_tile = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t_done = 0; int n_tr = 0;
    for (int b = 0; b < batches; ++b) {
      long long t0 = clock64();
      uint32_t accumulate = 0;
      if (v.mode == 1) {
        // whole GEMM1 inside ONE elected region, runtime loops
        if (elect_one()) {
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < v.taps / 3; ++kw) {
              const uint64_t a_desc = hi | ((a_tile + (kh * v.wp + kw) * v.swb) >> 4);
              const uint64_t b_desc = hi | ((b0 + (kh * 3 + kw) * v.b_stride) >> 4);
              for (int ks = 0; ks < v.nks; ++ks) umma_i8(tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, accumulate | ks);
              accumulate = 1;
            }
        }
        __syncwarp();
      } else if (v.mode == 2) {
        // same, selected with lane == 0 instead of elect.sync
        if (lane == 0) {
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < v.taps / 3; ++kw) {
              const uint64_t a_desc = hi | ((a_tile + (kh * v.wp + kw) * v.swb) >> 4);
              const uint64_t b_desc = hi | ((b0 + (kh * 3 + kw) * v.b_stride) >> 4);
              for (int ks = 0; ks < v.nks; ++ks) umma_i8(tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, accumulate | ks);
              accumulate = 1;
            }
        }
        __syncwarp();
      } else if (v.mode == 3) {
        // per-tap elect, but NO __syncwarp between taps
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < v.taps / 3; ++kw) {
            const uint64_t a_desc = hi | ((a_tile + (kh * v.wp + kw) * v.swb) >> 4);
            const uint64_t b_desc = hi | ((b0 + (kh * 3 + kw) * v.b_stride) >> 4);
            if (elect_one()) {
              for (int ks = 0; ks < v.nks; ++ks) umma_i8(tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, accumulate | ks);
            }
            accumulate = 1;
          }
        __syncwarp();
      } else if (v.mode == 4) {
        // all 32 lanes execute the loop; the MMA itself is predicated on an elected lane computed ONCE
        const bool leader = elect_one();
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < v.taps / 3; ++kw) {
            const uint64_t a_desc = hi | ((a_tile + (kh * v.wp + kw) * v.swb) >> 4);
            const uint64_t b_desc = hi | ((b0 + (kh * 3 + kw) * v.b_stride) >> 4);
            for (int ks = 0; ks < v.nks; ++ks)
              if (leader) umma_i8(tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, accumulate | ks);
            accumulate = 1;
          }
        __syncwarp();
      } else
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < v.taps / 3; ++kw) {
          const uint32_t a_tap = a_tile + (kh * v.wp + kw) * v.swb;
          const uint32_t b_base = b0 + (kh * 3 + kw) * v.b_stride;
          const uint64_t a_desc = hi | (a_tap >> 4), b_desc = hi | (b_base >> 4);
          if (elect_one()) {
            for (int ks = 0; ks < v.nks; ++ks) umma_i8(tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, accumulate | ks);
          }
          __syncwarp();
          accumulate = 1;
          if (v.trace && lane == 0 && n_tr < 4096) tr[n_tr++] = clock64();
        }
      if (elect_one()) umma_commit(smem_u32(&bar_mma));
      __syncwarp();
      mbar_wait_warp(smem_u32(&bar_mma), b & 1);
      t_done += clock64() - t0;
    }
    if (lane == 0) out_done[blockIdx.x] = t_done / batches;
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

static void run_variants() {
  long long* d_done; unsigned long long* tr;
  CK(cudaMalloc(&d_done, 148 * 8)); CK(cudaMalloc(&tr, 4096 * 8));
  size_t smem = 170 * 1024;
  CK(cudaFuncSetAttribute(kv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct { const char* name; V v; } cases[] = {
      {"per-tap elect + syncwarp (conv v1)", {2, 58, 64, 4096, 0, 9, 0}},
      {"one elect around the whole GEMM1", {2, 58, 64, 4096, 0, 9, 1}},
      {"lane==0 around the whole GEMM1", {2, 58, 64, 4096, 0, 9, 2}},
      {"per-tap elect, no syncwarp between taps", {2, 58, 64, 4096, 0, 9, 3}},
      {"leader elected once, predicated MMAs", {2, 58, 64, 4096, 0, 9, 4}},
      {"one elect, sw128 nks=4 (36 MMAs)", {4, 29, 128, 8192, 0, 9, 1}},
      {"leader once, sw128 nks=4 (36 MMAs)", {4, 29, 128, 8192, 0, 9, 4}},
  };
  for (auto& c : cases) {
    kv<<<148, 128, smem>>>(200, c.v, d_done, tr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant launch error %s\n", cudaGetErrorString(e)); return; }
    long long d; CK(cudaMemcpy(&d, d_done, 8, cudaMemcpyDeviceToHost));
    printf("variant %-46s: %6lld cycles per GEMM1 (%d MMAs, ideal %d)\n", c.name, d, 9 * c.v.nks, 9 * c.v.nks * 48);
  }
}

int main() {
  run_w();
  run_variants();
  long long *d_issue, *d_done; float* sink;
  int nsm = 148;
  CK(cudaMalloc(&d_issue, nsm * 8)); CK(cudaMalloc(&d_done, nsm * 8)); CK(cudaMalloc(&sink, 4096));
  size_t smem = (256 + 64) * 128 + 2048;
  CK(cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[] = {"park", "spin all lanes", "spin one lane", "FFMA loop", "LDS.128 loop", "tcgen05.ld loop", "spin one lane+nanosleep"};
  for (int companions : {0, 3, 11})
    for (int mma_last : {0, 1})
      for (int mode = 0; mode < 7; ++mode) {
        if (companions == 0 && (mode > 0 || mma_last)) continue;
        k<64><<<nsm, 32 * (1 + companions), smem>>>(200, mode, mma_last, d_issue, d_done, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch error %s (companions=%d mode=%d)\n", cudaGetErrorString(e), companions, mode); return 1; }
        std::vector<long long> a(nsm), b(nsm);
        CK(cudaMemcpy(a.data(), d_issue, nsm * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d_done, nsm * 8, cudaMemcpyDeviceToHost));
        printf("companions=%2d mma_warp=%-5s %-24s: issue %5lld cyc/batch(18 MMA)  done %5lld cyc/batch  (ideal 864)\n",
               companions, mma_last ? "last" : "first", names[mode], a[0], b[0]);
      }
  return 0;
}
