/* xbyak/xbyak_util.h -- NOT Xbyak: the CPU-feature part of the recording shim (see xbyak.h).
 * TEST INFRASTRUCTURE (oracle/_ref).  Xbyak::util::Cpu as used by src/jit_generator.h:45-143. */
#ifndef DF_XBYAK_SHIM_UTIL_H_
#define DF_XBYAK_SHIM_UTIL_H_
#include <stdlib.h>

namespace Xbyak {
namespace util {

class Cpu {
 public:
  enum Type {
    tSSE42, tAVX2, tAVX512F, tAVX512BW, tAVX512VL, tAVX512DQ, tAVX512_VNNI, tAVX512CD, tAVX512ER, tAVX512PF,
    tAVX512_4FMAPS, tAVX512_4VNNIW
  };
  static const unsigned int maxNumberCacheLevels = 10;
  unsigned int data_cache_levels;
  unsigned int data_cache_size[maxNumberCacheLevels];
  unsigned int cores_sharing_data_cache[maxNumberCacheLevels];

  Cpu() : data_cache_levels(0) {  /* 0: the reference then falls back to its own cache-size defaults */
    for (unsigned i = 0; i < maxNumberCacheLevels; ++i) data_cache_size[i] = cores_sharing_data_cache[i] = 0;
    __builtin_cpu_init();
  }
  bool has(Type t) const {
    switch (t) {
      case tSSE42: return __builtin_cpu_supports("sse4.2");
      case tAVX2: return __builtin_cpu_supports("avx2");
      case tAVX512F: return __builtin_cpu_supports("avx512f");
      case tAVX512BW: return __builtin_cpu_supports("avx512bw");
      case tAVX512VL: return __builtin_cpu_supports("avx512vl");
      case tAVX512DQ: return __builtin_cpu_supports("avx512dq");
      /* DFREF_NO_VNNI=1 makes the reference emit its vpmaddubsw / vpmaddwd / vpaddd fallback */
      case tAVX512_VNNI: return __builtin_cpu_supports("avx512vnni") && !getenv("DFREF_NO_VNNI");
      case tAVX512CD: return __builtin_cpu_supports("avx512cd");
      default: return false; /* Knights-family extensions */
    }
  }
};

}  // namespace util
}  // namespace Xbyak
#endif
