// tma_store_probe.cu -- what does a SWIZZLE_128B TMA tile STORE accept?
//   (a) shared-memory source that is 128 B- but not 1024 B-aligned (absolute-address swizzle?)
//   (b) negative start coordinates (box partly left of the tensor: clipped?)
//   (c) start coordinate > 0 with the box running past the tensor edge (clipped?)
// Each case is its own launch so that a rejected instruction is attributed to it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_store_probe tma_store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../deep-fusion_b200/csrc/sm100_ptx.cuh"
using namespace sm100;

constexpr int N = 2, H = 4, W = 28, C = 256, WP = 29;

__global__ void k(const __grid_constant__ CUtensorMap tm, int base_row, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  // 160 rows of 128 B; logical (row r, 16-byte unit u) lives at unit u ^ (r & 7)
  for (int i = threadIdx.x; i < 160 * 128; i += blockDim.x) {
    const int r = i >> 7, b = i & 127, u = b >> 4;
    sm[r * 128 + (((u ^ (r & 7)) << 4) | (b & 15))] = (uint8_t)(r * 7 + b * 3 + 1);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_4d(&tm, smem_u32(sm) + base_row * 128, c0, c1, c2, c3);
    bulk_commit_group();
    bulk_wait_all();
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  uint8_t* d;
  const size_t bytes = (size_t)N * H * W * C;
  cudaMalloc(&d, bytes);
  CUtensorMap tm;
  cuuint64_t gd[4] = {C, W, H, N};
  cuuint64_t gs[3] = {C, (cuuint64_t)W * C, (cuuint64_t)H * W * C};
  cuuint32_t box[4] = {128, WP, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  struct Case { const char* name; int base_row, c0, c1, c2, c3; } cases[] = {
      {"aligned base, col 0", 0, 128, 0, 1, 1},
      {"base row 5 (128 B aligned only), col 0", 5, 0, 0, 2, 0},
      {"col +12 (box runs past W)", 0, 128, 12, 3, 1},
      {"col -10 (box starts left of the tensor)", 16, 0, -10, 0, 1},
      {"base row 99, col -17", 99, 128, -17, 2, 1},
  };
  std::vector<uint8_t> h(bytes);
  for (const Case& c : cases) {
    cudaMemset(d, 0, bytes);
    k<<<1, 256, 32768>>>(tm, c.base_row, c.c0, c.c1, c.c2, c.c3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-45s: LAUNCH FAILED: %s\n", c.name, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(h.data(), d, bytes, cudaMemcpyDeviceToHost);
    long bad = 0, written = 0;
    for (int n = 0; n < N; ++n)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          for (int ch = 0; ch < C; ++ch) {
            const uint8_t got = h[(((size_t)n * H + y) * W + x) * C + ch];
            uint8_t want = 0;
            const int br = x - c.c1, bc = ch - c.c0;  // position inside the box
            if (n == c.c3 && y == c.c2 && br >= 0 && br < WP && bc >= 0 && bc < 128)
              want = (uint8_t)((c.base_row + br) * 7 + bc * 3 + 1);
            bad += got != want;
            written += got != 0;
          }
    printf("%-45s: %s (%ld bytes non-zero, %ld wrong)\n", c.name, bad ? "WRONG" : "ok", written, bad);
  }
  return 0;
}
