// Host-side throughput of the staged transfer's narrowing step (int32 -> uint16) on the box's cores, against plain reads and
// memcpy: decides how the end-to-end pipeline splits its bytes between the narrowed and the direct copies.
//   g++ -O3 -std=c++17 -pthread tools/host_bw.cpp -o /tmp/host_bw && /tmp/host_bw [GiB] [threads]
#include <immintrin.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class F>
static double run_threads(int nt, size_t n, F f) {
  const double t0 = now();
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { f(n * t / nt, n * (t + 1) / nt, t); });
  for (auto& x : th) x.join();
  return now() - t0;
}

static unsigned narrow_scalar(const int32_t* src, uint16_t* dst, size_t a, size_t b) {
  unsigned acc = 0;
  for (size_t i = a; i < b; ++i) { const uint32_t v = (uint32_t)src[i]; acc |= (v > 65534u) ? 1u : 0u; dst[i] = (uint16_t)v; }
  return acc;
}

__attribute__((target("avx2"))) static unsigned narrow_avx2(const int32_t* src, uint16_t* dst, size_t a, size_t b, bool stream) {
  size_t i = a;
  unsigned acc = 0;
  while (i < b && ((uintptr_t)(dst + i) & 31)) { const uint32_t v = (uint32_t)src[i]; acc |= v > 65534u; dst[i] = (uint16_t)v; ++i; }
  __m256i bad = _mm256_setzero_si256();
  const __m256i lim = _mm256_set1_epi32(65534);
  for (; i + 16 <= b; i += 16) {
    const __m256i x0 = _mm256_loadu_si256((const __m256i*)(src + i)), x1 = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    // unsigned v > 65534  <=>  max_u(v, 65535) == v with v >= 65535
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_cmpgt_epi32(_mm256_max_epu32(x0, lim), lim), _mm256_cmpgt_epi32(_mm256_max_epu32(x1, lim), lim)));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_srai_epi32(x0, 31), _mm256_srai_epi32(x1, 31)));      // negative values
    __m256i p = _mm256_packus_epi32(x0, x1);                   // lanes interleaved: [x0.lo x1.lo | x0.hi x1.hi]
    p = _mm256_permute4x64_epi64(p, 0xD8);
    if (stream) _mm256_stream_si256((__m256i*)(dst + i), p);
    else _mm256_storeu_si256((__m256i*)(dst + i), p);
  }
  acc |= _mm256_testz_si256(bad, bad) ? 0u : 1u;
  for (; i < b; ++i) { const uint32_t v = (uint32_t)src[i]; acc |= v > 65534u; dst[i] = (uint16_t)v; }
  if (stream) _mm_sfence();
  return acc;
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 2.0;
  const int max_t = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
  const size_t n = (size_t)(gib * (1u << 30)) / 4;
  int32_t* src = (int32_t*)aligned_alloc(64, n * 4);
  uint16_t* dst = (uint16_t*)aligned_alloc(64, n * 2);
  int32_t* cpy = (int32_t*)aligned_alloc(64, n * 4);
  run_threads(max_t, n, [&](size_t a, size_t b, int) { for (size_t i = a; i < b; ++i) { src[i] = (int32_t)(i % 97); dst[i] = 0; cpy[i] = 0; } });
  std::printf("buffer %.2f GiB int32, hardware threads %u, avx2 %d\n", gib, std::thread::hardware_concurrency(), (int)__builtin_cpu_supports("avx2"));
  std::vector<long long> sink(64);
  for (int nt : {1, 2, 4, 8, 16, 32}) {
    if (nt > max_t) break;
    double best[5] = {1e9, 1e9, 1e9, 1e9, 1e9};
    for (int rep = 0; rep < 3; ++rep) {
      best[0] = std::min(best[0], run_threads(nt, n, [&](size_t a, size_t b, int t) { long long s = 0; for (size_t i = a; i < b; ++i) s += src[i]; sink[t] = s; }));
      best[1] = std::min(best[1], run_threads(nt, n, [&](size_t a, size_t b, int) { std::memcpy(cpy + a, src + a, (b - a) * 4); }));
      best[2] = std::min(best[2], run_threads(nt, n, [&](size_t a, size_t b, int t) { sink[t] = narrow_scalar(src, dst, a, b); }));
      if (__builtin_cpu_supports("avx2")) {
        best[3] = std::min(best[3], run_threads(nt, n, [&](size_t a, size_t b, int t) { sink[t] = narrow_avx2(src, dst, a, b, false); }));
        best[4] = std::min(best[4], run_threads(nt, n, [&](size_t a, size_t b, int t) { sink[t] = narrow_avx2(src, dst, a, b, true); }));
      }
    }
    const double gb = n * 4 / 1e9;
    std::printf("threads %2d | read %6.1f GB/s | memcpy %6.1f GB/s (src bytes) | narrow scalar %6.1f | avx2 %6.1f | avx2 + streaming stores %6.1f  (GB/s of int32 consumed)\n",
                nt, gb / best[0], gb / best[1], gb / best[2], gb / best[3], gb / best[4]);
  }
  return (int)(sink[0] & 1);
}
