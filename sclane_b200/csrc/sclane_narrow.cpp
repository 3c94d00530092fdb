// Host side of the staged transfer (include/sclane_b200.h, sclane_opts.h2d_mode): int32 counts -> uint16, one contiguous range per
// call (the caller spreads ranges over host threads).  A plain C++ translation unit -- compiled by the host compiler, not parsed
// by the CUDA front end -- so that the AVX2 variant can be selected at run time.
//
// Replaces nothing in the reference: testDynamic.R:145-147 writes the dense integer matrix to a file-backed matrix and every
// worker reads its gene column from there; here the same int32 bytes are halved on their way to the device because the PCIe
// link (55 GB/s), not the GPU, bounds the end-to-end time of a 16 GB matrix.
#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define SCLANE_X86 1
#else
#define SCLANE_X86 0
#endif

namespace sclane {

// false when a value does not fit (negative or > 65534: 65535 is kept free as a guard value)
static bool narrow_range_scalar(const int32_t* src, size_t n, uint16_t* dst) {
  unsigned acc = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t v = (uint32_t)src[i];
    acc |= (v > 65534u) ? 1u : 0u;
    dst[i] = (uint16_t)v;
  }
  return acc == 0;
}

#if SCLANE_X86
// 16 values per iteration; non-temporal stores: the staging buffer is written once and read by the DMA engine only, so
// keeping it out of the caches saves the read-for-ownership traffic (a third of the memory traffic of the step).
__attribute__((target("avx2"))) static bool narrow_range_avx2(const int32_t* src, size_t n, uint16_t* dst) {
  size_t i = 0;
  unsigned acc = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) { const uint32_t v = (uint32_t)src[i]; acc |= v > 65534u; dst[i] = (uint16_t)v; ++i; }
  __m256i bad = _mm256_setzero_si256();
  const __m256i lim = _mm256_set1_epi32(65534);
  for (; i + 16 <= n; i += 16) {
    const __m256i x0 = _mm256_loadu_si256((const __m256i*)(src + i)), x1 = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    // as unsigned: v > 65534  <=>  max(v, 65534) > 65534 as signed, or v negative as signed
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_cmpgt_epi32(_mm256_max_epu32(x0, lim), lim), _mm256_cmpgt_epi32(_mm256_max_epu32(x1, lim), lim)));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_srai_epi32(x0, 31), _mm256_srai_epi32(x1, 31)));
    __m256i p = _mm256_packus_epi32(x0, x1);            // per 128-bit lane: [x0.lo x1.lo | x0.hi x1.hi]
    p = _mm256_permute4x64_epi64(p, 0xD8);               // -> x0 then x1
    _mm256_stream_si256((__m256i*)(dst + i), p);
  }
  acc |= _mm256_testz_si256(bad, bad) ? 0u : 1u;
  for (; i < n; ++i) { const uint32_t v = (uint32_t)src[i]; acc |= v > 65534u; dst[i] = (uint16_t)v; }
  _mm_sfence();
  return acc == 0;
}
#endif

bool narrow_range_u16(const int32_t* src, size_t n, uint16_t* dst) {
#if SCLANE_X86
  static const bool has_avx2 = __builtin_cpu_supports("avx2");
  if (has_avx2) return narrow_range_avx2(src, n, dst);
#endif
  return narrow_range_scalar(src, n, dst);
}

}  // namespace sclane
