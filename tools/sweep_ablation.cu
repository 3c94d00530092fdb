// sweep_ablation.cu -- the product k_sweep_moments kernel with parts switched off at compile time, to see what each
// costs on top of the bare access pattern (tools/stream_probe.cu).  Built several times by tools/sweep_ablation.sh.
#include <cstdio>
#include <vector>
#include "../sclane_b200/csrc/sclane_device.cuh"
using namespace sclane;
int main(int argc, char** argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 2000, N = argc > 2 ? atoi(argv[2]) : 10000, T = 25;
  Lineage L; memset(&L, 0, sizeof(L));
  L.N = N; L.Npad = (N + 127) / 128 * 128; L.T = T;
  for (int b = 0; b <= T + 1; ++b) L.bstart[b] = (int)((long long)N * b / (T + 1));
  for (int b = T + 2; b <= NBMAX; ++b) L.bstart[b] = N;
  Lineage* dL; int* ys; double *mu, *u, *mom; GeneOutDev* out; char* flush; const size_t fn = (size_t)256 << 20;
  cudaMalloc(&dL, sizeof(L)); cudaMemcpy(dL, &L, sizeof(L), cudaMemcpyHostToDevice);
  cudaMalloc(&ys, (size_t)G * L.Npad * 4); cudaMalloc(&mu, (size_t)G * L.Npad * 8); cudaMalloc(&u, (size_t)L.Npad * 8);
  cudaMalloc(&mom, (size_t)G * NBMAX * NACC * 8); cudaMalloc(&out, (size_t)G * sizeof(GeneOutDev)); cudaMalloc(&flush, fn);
  cudaMemset(ys, 0, (size_t)G * L.Npad * 4); cudaMemset(mu, 0, (size_t)G * L.Npad * 8); cudaMemset(u, 0, (size_t)L.Npad * 8);
  cudaMemset(out, 0, (size_t)G * sizeof(GeneOutDev));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 8; ++rep) {
    cudaMemsetAsync(flush, rep, fn);
    cudaEventRecord(e0);
    k_sweep_moments<false><<<G, kSweepThreads>>>(dL, ys, mu, u, out, mom, G);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("cell2 %d warps %d minb %d prefetch %d flush %d boundary %d epilogue %d : %.1f us  %.0f GB/s  (%s)\n", SCLANE_SWEEP_CELL2, SCLANE_SWEEP_WARPS, SCLANE_SWEEP_MINB,
         SCLANE_SWEEP_PREFETCH, SCLANE_SWEEP_DO_FLUSH, SCLANE_SWEEP_DO_BOUNDARY, SCLANE_SWEEP_DO_EPILOGUE, best * 1e3, 12.0 * N * G / best / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
