// FP64 peak probe for B200 (sm_100a): DMMA m8n8k4 issue rate vs DFMA issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double *out, int iters) {
  double c0[ILP], c1[ILP];
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters) {
  double c[ILP];
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double *out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int rep = 0; rep < 2; rep++) {
      int blocks = p.multiProcessorCount * (1024 / threads) ;
      cudaEventRecord(e0);
      k_dmma<8><<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 8 * 8 * 4 * 8 * (double)iters * (threads / 32) * blocks;
      if (rep) printf("DMMA threads %4d blocks %d: %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, flops / ms * 1e-9);
      cudaEventRecord(e0);
      k_dfma<8><<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      flops = 2.0 * 8 * (double)iters * threads * blocks;
      if (rep) printf("DFMA threads %4d blocks %d: %.3f ms  %.2f TFLOP/s\n", threads, blocks, ms, flops / ms * 1e-9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
