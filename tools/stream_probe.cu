// stream_probe.cu -- access-pattern floor for the score-sweep streaming kernel: same loads (int4 counts, 2x double2
// means, 2x double2 shared coordinates), trivial math, no bucket logic.  Build/run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/stream_probe.cu -o /tmp/stream_probe && /tmp/stream_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ int4 ld_i4(const int* p) { int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ double2 ld_d2(const double* p) { double2 r; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r; }

// MODE 0: warp = contiguous tile range (as k_sweep_moments); MODE 1: tiles interleaved across the CTA's warps
template <int WARPS, int MODE, int MATH>
__global__ void probe(const int* ys, const double* mu, const double* u, double* out, int N, int Npad, int G, double sigma) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    const int* y = ys + (size_t)g * Npad;
    const double* m = mu + (size_t)g * Npad;
    const int ntiles = Npad >> 7;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    int t0, t1, ts;
    if (MODE == 0) { t0 = ntiles * warp / WARPS; t1 = ntiles * (warp + 1) / WARPS; ts = 1; }
    else { t0 = warp; t1 = ntiles; ts = WARPS; }
    for (int t = t0; t < t1; t += ts) {
      const int i0 = (t << 7) + (lane << 2);
      const int4 y4 = ld_i4(y + i0);
      const double2 ma = ld_d2(m + i0), mb = ld_d2(m + i0 + 2);
      const double2 ua = __ldg((const double2*)(u + i0)), ub = __ldg((const double2*)(u + i0 + 2));
      const int yv[4] = {y4.x, y4.y, y4.z, y4.w};
      const double mv[4] = {ma.x, ma.y, mb.x, mb.y}, uv[4] = {ua.x, ua.y, ub.x, ub.y};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MATH) {
          const double yd = (double)yv[j];
          const double inv = 1.0 / fma(sigma, mv[j], 1.0);
          const double w = mv[j] * inv, r = (yd - mv[j]) * inv, wu = w * uv[j];
          a0 += w; a1 += wu; a2 = fma(wu, uv[j], a2); a3 += r; a4 = fma(r, uv[j], a4);
        } else { a0 += mv[j]; a1 += uv[j]; a2 += (double)yv[j]; }
      }
    }
    double s = a0 + a1 + a2 + a3 + a4;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(out + g, s);
  }
}

template <int WARPS, int MODE, int MATH>
void run(const char* name, int grid, const int* ys, const double* mu, const double* u, double* out, int N, int Npad, int G, char* flush, size_t flush_n) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaMemsetAsync(flush, rep, flush_n);
    cudaEventRecord(e0);
    probe<WARPS, MODE, MATH><<<grid, WARPS * 32>>>(ys, mu, u, out, N, Npad, G, 0.3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double bytes = 12.0 * (double)N * G;
  printf("%-44s grid %5d  %.1f us  %.0f GB/s\n", name, grid, best * 1e3, bytes / best / 1e6);
}

int main(int argc, char** argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 2000, N = argc > 2 ? atoi(argv[2]) : 10000;
  const int Npad = (N + 127) / 128 * 128;
  int* ys; double *mu, *u, *out; char* flush; const size_t flush_n = (size_t)256 << 20;
  cudaMalloc(&ys, (size_t)G * Npad * 4); cudaMalloc(&mu, (size_t)G * Npad * 8); cudaMalloc(&u, (size_t)Npad * 8);
  cudaMalloc(&out, (size_t)G * 8); cudaMalloc(&flush, flush_n);
  cudaMemset(ys, 0, (size_t)G * Npad * 4); cudaMemset(mu, 0, (size_t)G * Npad * 8); cudaMemset(u, 0, (size_t)Npad * 8); cudaMemset(out, 0, (size_t)G * 8);
  printf("G %d N %d: %.1f MB per launch\n", G, N, 12.0 * N * G / 1e6);
  run<8, 0, 0>("8 warps contiguous, no math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<8, 1, 0>("8 warps interleaved, no math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<8, 0, 1>("8 warps contiguous, math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<8, 1, 1>("8 warps interleaved, math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<4, 0, 1>("4 warps contiguous, math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<4, 1, 1>("4 warps interleaved, math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<8, 1, 1>("8 warps interleaved, math, persistent 148x4", 148 * 4, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<8, 0, 1>("8 warps contiguous, math, persistent 148x4", 148 * 4, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<8, 1, 1>("8 warps interleaved, math, persistent 148x8", 148 * 8, ys, mu, u, out, N, Npad, G, flush, flush_n);
  run<16, 1, 1>("16 warps interleaved, math, grid=G", G, ys, mu, u, out, N, Npad, G, flush, flush_n);
  return 0;
}
