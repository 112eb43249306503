// Inner-loop probe for the trailing update: operands resident in shared memory (same pitches and fragment
// layout as front_update_kernel), no global traffic.  What does the smem -> DMMA loop sustain by itself?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_core_probe dmma_core_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int K = 64, UROW = 132, UCOL = 68;

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: fragments from smem each k-step (4x4 blocking: 8 LDS.64 per 16 DMMA)
// MODE 1: fragments loaded once, DMMA only (register operands)
// MODE 2: 4x4 blocking, explicit double buffering of the fragments
template <int MODE>
__global__ void __launch_bounds__(256) k_core(double *out, int tiles) {
  extern __shared__ double smem[];
  double *sA = smem, *sB = smem + K * UROW;
  for (int i = threadIdx.x; i < K * (UROW + UCOL); i += blockDim.x) smem[i] = 1e-3 * (i % 17);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = (warp >> 2) & 1;
  const int g = lane >> 2, q = lane & 3;
  const double *pa = sA + q * UROW + wm * 32 + g;
  const double *pb = sB + q * UCOL + wn * 32 + g;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int t = 0; t < tiles; t++) {
    if (MODE == 1) {
      double a[4], b[4];
#pragma unroll
      for (int f = 0; f < 4; f++) { a[f] = pa[f * 8]; b[f] = pb[f * 8]; }
      for (int k = 0; k < K; k += 4) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else if (MODE == 0) {
      for (int k = 0; k < K; k += 4) {
        double a[4], b[4];
#pragma unroll
        for (int f = 0; f < 4; f++) { a[f] = pa[k * UROW + f * 8]; b[f] = pb[k * UCOL + f * 8]; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else {
      double a0[4], b0[4], a1[4], b1[4];
#pragma unroll
      for (int f = 0; f < 4; f++) { a0[f] = pa[f * 8]; b0[f] = pb[f * 8]; }
      for (int k = 0; k < K; k += 8) {
#pragma unroll
        for (int f = 0; f < 4; f++) { a1[f] = pa[(k + 4) * UROW + f * 8]; b1[f] = pb[(k + 4) * UCOL + f * 8]; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a0[i], b0[j]);
        if (k + 8 < K) {
#pragma unroll
          for (int f = 0; f < 4; f++) { a0[f] = pa[(k + 8) * UROW + f * 8]; b0[f] = pb[(k + 8) * UCOL + f * 8]; }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a1[i], b1[j]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double *out, int ctas_per_sm, int threads, int sms) {
  const size_t smem = (size_t)K * (UROW + UCOL) * sizeof(double);
  cudaFuncSetAttribute(k_core<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int tiles = 2000;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    k_core<MODE><<<sms * ctas_per_sm, threads, smem>>>(out, tiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 512.0 * 16 * (K / 4) * (double)tiles * (threads / 32) * sms * ctas_per_sm;
    if (rep) printf("%-28s ctas/SM %d threads %4d: %8.3f ms  %6.2f TFLOP/s\n", name, ctas_per_sm, threads, ms, flops / ms * 1e-9);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *out; cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  for (int c : {1, 2}) {
    run<1>("register operands", out, c, 256, sms);
    run<0>("smem fragments", out, c, 256, sms);
    run<2>("smem fragments, dbl-buffered", out, c, 256, sms);
  }
  run<1>("register operands", out, 1, 128, sms);
  run<0>("smem fragments", out, 1, 128, sms);
  run<0>("smem fragments", out, 4, 128, sms);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
