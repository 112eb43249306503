// One-shot all-reduce of a SMALL buffer over peer memory (NVLink / NVSwitch), fused with the cross-rank handshake.
//
// The Schur-complement path has exactly one exchange step per phase (reference
// mpi_explicit_schur_complement.py:343 `Allreduce` of the S values, :387 of the coupling right-hand side), and at
// BASELINE config 2 the payloads are 20 kB and 400 B: NCCL's latency (50-60 us at 8 ranks) is then
// several times the transfer.  Here every rank's contribution lives in a buffer that all ranks have mapped
// (torch symmetric memory); ONE kernel per rank raises its flag in every peer's signal pad, waits for the peers'
// flags, reads all contributions straight out of the peers' memory and adds them in RANK ORDER -- so every rank
// gets bit-identical sums (NCCL's choice of algorithm does not promise that) -- into a local output buffer.
// Contributions are double-buffered by the caller (a buffer is rewritten two exchanges later, when every peer has
// passed the next handshake), so no second handshake is needed.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ppb {

constexpr int PEER_MAX = 16;

struct PeerPtrs {
  const double *buf[PEER_MAX];   // rank q's contribution (mapped in this process)
  uint32_t *sig[PEER_MAX];       // rank q's signal pad
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer(const double *p) {  // never from a stale L1 line: the address is reused
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// flags of channel `slot`: sig[q][slot + r] = last sequence number rank r announced to rank q.  A few CTAs: the first
// one announces, every one waits for the peers' flags itself (no grid barrier) and sums its share of the elements.
__global__ void __launch_bounds__(1024) peer_allreduce_kernel(PeerPtrs P, int rank, int world, int slot, uint32_t seq,
                                                              int n, double *__restrict__ out) {
  if ((int)threadIdx.x < world) {
    const int q = threadIdx.x;
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(P.sig[q] + slot + rank, seq);          // "my contribution number seq is complete"
    }
    const uint32_t *mine = P.sig[rank] + slot + q;
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) { }    // rank q's contribution is complete (wrap-safe)
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v[PEER_MAX];
#pragma unroll
    for (int q = 0; q < PEER_MAX; ++q) v[q] = q < world ? ld_peer(P.buf[q] + i) : 0.0;   // all loads in flight
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < PEER_MAX; ++q)
      if (q < world) s += v[q];                                                         // rank order
    out[i] = s;
  }
}

}  // namespace ppb
