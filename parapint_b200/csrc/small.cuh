// Whole-front factorisation in shared memory for SMALL dense fronts (block roots with narrow borders and small
// coupling matrices): one CTA loads the active rows of the front (pivot candidates + border rows, skipping unused
// delayed-pivot slots), eliminates every pivot column with the same threshold-pivoting routine the multifrontal
// subtree uses (1x1 / 2x2 pivots chosen and tested among the first n rows only, as the reference factorises K_i
// without its border), and writes L, D, the pivot flags, the
// permutation and the trailing block (-A K^-1 A^T) back in the layout the batched solve kernels expect.
// One launch replaces the (panel, interchange, update) launch sequence of factor.cuh when every front of the
// batch fits; each pivot step then costs shared-memory latency instead of a round trip to L2.
#pragma once
#include "sparse.cuh"

namespace ppb {

constexpr int SM_CAP = 164, SM_LD = SM_CAP + 1;   // 164 x 165 doubles = 211 KB of the 227 KB a CTA may use
constexpr size_t SM_SMEM = fb_bytes(SM_CAP, SM_LD) + 16;

__global__ void __launch_bounds__(SF_NT) front_small_kernel(const Front *__restrict__ fronts, double u,
                                                            double pivtol) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ int cnt[3];
  const Front F = fronts[blockIdx.x];
  const int tid = threadIdx.x;
  const int n = F.n, m = F.m, nb = F.nb, S = n + m;
  if (n == 0 || F.state[ST_KCUR] >= n) return;
  if (S > SM_CAP) {  // excluded by the host-side classification (static n + m <= SM_CAP)
    if (tid == 0 && F.state[ST_INFO] == 0) F.state[ST_INFO] = 1;
    return;
  }
  const FrontBuf B = carve(sm_raw, SM_CAP, SM_LD);
  const int ldA = F.ld;
  double *__restrict__ A = F.A;
  for (int idx = tid; idx < S * S; idx += SF_NT) {
    const int j = idx / S, i = idx - j * S;
    if (i < j) continue;
    const int si = i < n ? i : nb + (i - n), sj = j < n ? j : nb + (j - n);
    B.F[i + j * SM_LD] = A[si + (size_t)sj * ldA];
  }
  for (int i = tid; i < S; i += SF_NT) {
    B.fid[i] = i;
    B.opos[i] = i;
    B.bsz[i] = 1;
    B.map[i] = i < n ? F.perm[i] : 0;
  }
  if (tid < 3) cnt[tid] = 0;
  __syncthreads();
  const int t = factor_front<SF_NT>(B, S, n, u, pivtol, cnt, n);
  __syncthreads();
  // columns that found no pivot: the remaining fully-summed block is (numerically) null -> singular
  for (int k = t + tid; k < n; k += SF_NT) {
    B.bsz[k] = 1;
    B.F[k + k * SM_LD] = 0.0;
  }
  __syncthreads();
  for (int idx = tid; idx < S * S; idx += SF_NT) {
    const int j = idx / S, i = idx - j * S;
    if (i < j) continue;
    const int si = i < n ? i : nb + (i - n), sj = j < n ? j : nb + (j - n);
    A[si + (size_t)sj * ldA] = B.F[i + j * SM_LD];
  }
  for (int k = tid; k < n; k += SF_NT) {
    F.bsz[k] = B.bsz[k];
    F.perm[k] = B.map[B.fid[k]];
  }
  if (tid == 0) {
    F.state[ST_KPREV] = n;
    F.state[ST_KCUR] = n;
    if (t < n && F.state[ST_INFO] == 0) F.state[ST_INFO] = t + 1;
  }
}

}  // namespace ppb
