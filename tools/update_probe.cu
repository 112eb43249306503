// Where does the trailing update spend its time?  Synthetic fronts of the config-5 root shape, one mid-factorisation
// launch, the production kernels with parts switched off (PP_UPDATE_PROBE hooks in factor.cuh).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPP_UPDATE_PROBE -I../include -o update_probe update_probe.cu
#include <cstdio>
#include <vector>
#include "../parapint_b200/csrc/factor.cuh"
using namespace ppb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
  const int NF = argc > 1 ? atoi(argv[1]) : 32;
  const int nf = 4707, n = 2082, nb = 2707, ld = 4720, kprev = 960, kcur = 1024;
  std::vector<Front> hf(NF);
  double *A, *W; int *state;
  CK(cudaMalloc(&A, (size_t)NF * ld * nf * sizeof(double)));
  CK(cudaMalloc(&W, (size_t)NF * ld * NBMAX * sizeof(double)));
  CK(cudaMalloc(&state, NF * 4 * sizeof(int)));
  CK(cudaMemset(A, 0, (size_t)NF * ld * nf * sizeof(double)));
  CK(cudaMemset(W, 0, (size_t)NF * ld * NBMAX * sizeof(double)));
  std::vector<int> hs(NF * 4, 0);
  for (int f = 0; f < NF; ++f) {
    hs[4 * f + ST_KCUR] = kcur; hs[4 * f + ST_KPREV] = kprev;
    Front &F = hf[f];
    F.A = A + (size_t)f * ld * nf; F.W = W + (size_t)f * ld * NBMAX; F.state = state + 4 * f;
    F.n = n; F.m = nf - nb; F.nf = nf; F.ld = ld; F.nb = nb; F.pad = 0;
    F.zbuf = F.bvec = nullptr; F.ipiv = F.bsz = F.perm = nullptr;
  }
  CK(cudaMemcpy(state, hs.data(), hs.size() * sizeof(int), cudaMemcpyHostToDevice));
  Front *df; CK(cudaMalloc(&df, NF * sizeof(Front)));
  CK(cudaMemcpy(df, hf.data(), NF * sizeof(Front), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(front_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM));
  CK(cudaFuncSetAttribute(front_update_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPS_SMEM));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // useful flops: compacted trailing order r = nf - (nb - n) - kcur, r^2 * kw
  const double r = nf - (nb - n) - kcur, flops = r * r * (kcur - kprev) * NF;
  auto report = [&](const char *name, float ms) { printf("%-44s %8.3f ms  %6.2f TFLOP/s useful\n", name, ms, flops / ms * 1e-9); };
  for (int probe : {0, 1, 2, 3, 8, 11}) {
    CK(cudaMemcpyToSymbol(g_update_probe, &probe, sizeof(int)));
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      front_update_kernel<<<dim3(update_tile_count(nf, kcur), NF), UPD_THREADS, UPD_SMEM>>>(df);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    char name[96];
    snprintf(name, sizeof name, "one tile per CTA%s%s%s", probe & 1 ? " -Cload" : "", probe & 2 ? " -Cstore" : "", probe & 8 ? " -kloop" : "");
    report(name, ms);
  }
  for (int S : {8, 12, 16, 24}) {
    for (int probe : {0, 3, 8}) {
      if (probe) continue;
      CK(cudaMemcpyToSymbol(g_update_probe, &probe, sizeof(int)));
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        front_update_strip_kernel<<<dim3(update_strip_count(nf, kcur, S), NF), UPD_THREADS, UPS_SMEM>>>(df, S);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
      }
      char name[96];
      snprintf(name, sizeof name, "strips S=%d%s%s%s%s", S, probe & 1 ? " -Cload" : "", probe & 2 ? " -Cstore" : "",
               probe & 4 ? " -Bstage" : "", probe & 8 ? " -kloop" : "");
      report(name, ms);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
