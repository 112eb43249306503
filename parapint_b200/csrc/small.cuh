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

// `add` (optional, one front only): an m_c x m_c column-major matrix added to the front while it is loaded -- the reduced
// Schur sum joining Q in the coupling front -- and `inertia_out` (optional) receives the pivot signs of the front, so
// that the coupling phase S = Q + sum, LDL^T, inertia is ONE launch instead of four.
__global__ void __launch_bounds__(SF_NT) front_small_kernel(const Front *__restrict__ fronts, double u,
                                                            double pivtol, const double *__restrict__ add = nullptr,
                                                            int add_ld = 0, unsigned long long *inertia_out = nullptr) {
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
  PP_TR(1999 + (gridDim.x > 1 ? 0 : 8) * 0);
  const FrontBuf B = carve(sm_raw, SM_CAP, SM_LD);
  const int ldA = F.ld;
  double *__restrict__ A = F.A;
  for (int idx = tid; idx < S * S; idx += SF_NT) {
    const int j = idx / S, i = idx - j * S;
    if (i < j) continue;
    const int si = i < n ? i : nb + (i - n), sj = j < n ? j : nb + (j - n);
    double v = A[si + (size_t)sj * ldA];
    if (add) v += add[si + (size_t)sj * add_ld];
    B.F[i + j * SM_LD] = v;
  }
  for (int i = tid; i < S; i += SF_NT) {
    B.fid[i] = i;
    B.opos[i] = i;
    B.bsz[i] = 1;
    B.map[i] = i < n ? F.perm[i] : 0;
  }
  if (tid < 3) cnt[tid] = 0;
  PP_TR(2000 + (gridDim.x > 1 ? 0 : 8));
  __syncthreads();
  PP_TR(2001 + (gridDim.x > 1 ? 0 : 8));
  // Pivots are chosen and tested among the first n rows only, so the n x n pivot block can be factorised on its own
  // (a quarter of the per-column update work when the border is as wide as the block) and the border rows brought
  // up afterwards without any per-column barrier: one warp per border row carries the row in registers through the
  // forward substitution, then one pass forms the trailing block.  Needs n <= 64 (two columns per lane) and
  // scratch for W = L_B D behind the front in shared memory; otherwise everything is done in one sweep.
  const bool two_phase = m > 0 && n <= 64 && (long)m * n <= (long)(SM_CAP - S) * SM_LD;
  const int t = factor_front<SF_NT>(B, two_phase ? n : S, n, u, pivtol, cnt, n);
  __syncthreads();
  PP_TR(2002 + (gridDim.x > 1 ? 0 : 8));
  if (two_phase) {
    double *W = B.F + (size_t)S * SM_LD;  // m x n, column-major: W[r + k * m]
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWS = SF_NT / 32, RW = 4;  // RW border rows per warp at a time: independent chains overlap
    const int k0 = lane, k1 = lane + 32;
    // pivot kinds as two 32-bit masks (uniform tests instead of a dependent shared-memory load per step)
    const unsigned two_lo = __ballot_sync(0xffffffffu, k0 < t && B.bsz[k0] == 2);
    const unsigned two_hi = __ballot_sync(0xffffffffu, k1 < t && B.bsz[k1] == 2);
    const int f0 = k0 < n ? B.fid[k0] : 0, f1 = k1 < n ? B.fid[k1] : 0;
    for (int rb = warp; rb < m; rb += NWS * RW) {
      double x0[RW], x1[RW];
#pragma unroll
      for (int q = 0; q < RW; ++q) {
        const int r = rb + q * NWS;
        // the row in pivot order (the pivot block's interchanges did not touch the border rows)
        x0[q] = (r < m && k0 < n) ? B.F[(n + r) + f0 * SM_LD] : 0.0;
        x1[q] = (r < m && k1 < n) ? B.F[(n + r) + f1 * SM_LD] : 0.0;
      }
      for (int j = 0; j < t;) {
        const bool pair = j < 32 ? (two_lo >> j) & 1u : (two_hi >> (j - 32)) & 1u;
        if (pair) {
          const int j1 = j + 1;
          const double la0 = (k0 > j1 && k0 < n) ? B.F[k0 + j * SM_LD] : 0.0, lb0 = (k0 > j1 && k0 < n) ? B.F[k0 + j1 * SM_LD] : 0.0;
          const double la1 = (k1 > j1 && k1 < n) ? B.F[k1 + j * SM_LD] : 0.0, lb1 = (k1 > j1 && k1 < n) ? B.F[k1 + j1 * SM_LD] : 0.0;
#pragma unroll
          for (int q = 0; q < RW; ++q) {
            const double w0 = __shfl_sync(0xffffffffu, j < 32 ? x0[q] : x1[q], j & 31);
            const double w1 = __shfl_sync(0xffffffffu, j1 < 32 ? x0[q] : x1[q], j1 & 31);
            x0[q] -= w0 * la0 + w1 * lb0;
            x1[q] -= w0 * la1 + w1 * lb1;
          }
          j += 2;
        } else {
          const double l0 = (k0 > j && k0 < n) ? B.F[k0 + j * SM_LD] : 0.0;
          const double l1 = (k1 > j && k1 < n) ? B.F[k1 + j * SM_LD] : 0.0;
#pragma unroll
          for (int q = 0; q < RW; ++q) {
            const double w0 = __shfl_sync(0xffffffffu, j < 32 ? x0[q] : x1[q], j & 31);
            x0[q] -= w0 * l0;
            x1[q] -= w0 * l1;
          }
          j += 1;
        }
      }
      __syncwarp();  // every lane has read its original entries of the rows before any is overwritten
      // x = W (rows of L_B D); L_B = W D^-1, pivot by pivot
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = h ? k1 : k0;
        const int bk = k < t ? B.bsz[k] : 1;
        // partner of a 2x2 pivot: the next column for its first member, the previous one for its second
        const int kp = bk == 2 ? k + 1 : (bk == 0 ? k - 1 : k);
        double s1 = 1.0, s2 = 0.0;  // l = s1 * x + s2 * partner
        if (k < t) {
          if (bk == 1) {
            s1 = 1.0 / B.F[k + k * SM_LD];
          } else {
            const int a = bk == 2 ? k : k - 1;  // first column of the pair
            const double e11 = B.F[a + a * SM_LD], e21 = B.F[a + 1 + a * SM_LD], e22 = B.F[a + 1 + (a + 1) * SM_LD];
            const double d11 = e22 / e21, d22 = e11 / e21;
            const double sc = (1.0 / (d11 * d22 - 1.0)) / e21;
            s1 = sc * (bk == 2 ? d11 : d22);
            s2 = -sc;
          }
        }
#pragma unroll
        for (int q = 0; q < RW; ++q) {
          const int r = rb + q * NWS;
          const double xv = h ? x1[q] : x0[q];
          const double p0 = __shfl_sync(0xffffffffu, x0[q], kp & 31), p1 = __shfl_sync(0xffffffffu, x1[q], kp & 31);
          const double xp = (kp & 32) ? p1 : p0;
          if (r < m && k < n) {
            W[r + k * m] = k < t ? xv : 0.0;
            B.F[(n + r) + k * SM_LD] = k < t ? s1 * xv + s2 * xp : xv;
          }
        }
      }
    }
    __syncthreads();
    PP_TR(2003);
    // trailing block: T -= L_B W^T (lower triangle), 2 x 2 outputs per thread
    const int mh = (m + 1) / 2;
    for (int idx = tid; idx < mh * mh; idx += SF_NT) {
      const int cb = idx / mh, rbk = idx - cb * mh;
      if (rbk < cb) continue;
      const int r = 2 * rbk, c = 2 * cb;
      const bool r1 = r + 1 < m, c1 = c + 1 < m;
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
      for (int k = 0; k < t; ++k) {
        const double l0 = B.F[(n + r) + k * SM_LD], l1 = r1 ? B.F[(n + r + 1) + k * SM_LD] : 0.0;
        const double w0 = W[c + k * m], w1 = c1 ? W[c + 1 + k * m] : 0.0;
        a00 += l0 * w0;
        a01 += l0 * w1;
        a10 += l1 * w0;
        a11 += l1 * w1;
      }
      B.F[(n + r) + (n + c) * SM_LD] -= a00;
      if (r1) B.F[(n + r + 1) + (n + c) * SM_LD] -= a10;
      if (c1 && r >= c + 1) B.F[(n + r) + (n + c + 1) * SM_LD] -= a01;
      if (r1 && c1) B.F[(n + r + 1) + (n + c + 1) * SM_LD] -= a11;
    }
    __syncthreads();
    PP_TR(2004);
  }
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
  PP_TR(2005 + (gridDim.x > 1 ? 0 : 8));
  if (tid == 0) {
    F.state[ST_KPREV] = n;
    F.state[ST_KCUR] = n;
    if (t < n && F.state[ST_INFO] == 0) F.state[ST_INFO] = t + 1;
  }
  if (inertia_out && tid < 3)   // pivots counted by factor_front; columns that found no pivot are zeros
    inertia_out[tid] = (unsigned long long)(cnt[tid] + (tid == 2 ? n - t : 0));
}

}  // namespace ppb
