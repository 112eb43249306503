// Batched symmetric-indefinite LDL^T of fronts: Bunch-Kaufman panel + FP64 tensor-core update.
//
// One launch of front_panel_kernel eliminates up to NB columns of every front (one CTA per front),
// choosing 1x1 / 2x2 pivots by the Bunch-Kaufman partial-pivoting rule among the first n rows.
// Columns needed by the pivot search are brought up to date lazily against the panel computed so
// far (W = L*D), so the O(n^2 NB) trailing work is left to front_update_kernel, a DMMA GEMM.
// Row interchanges are applied to the whole row of L (front_swaps_left_kernel), i.e. the factors
// satisfy P K P^T = L D L^T with one permutation, which keeps the triangular solves plain.
#pragma once
#include <cooperative_groups.h>
#include "front.cuh"

namespace ppb {

template <int NT>
__device__ __forceinline__ void block_argmax(double &val, int &idx, double *sval, int *sidx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    double v2 = __shfl_down_sync(0xffffffffu, val, o);
    int i2 = __shfl_down_sync(0xffffffffu, idx, o);
    if (i2 >= 0 && (idx < 0 || v2 > val || (v2 == val && i2 < idx))) { val = v2; idx = i2; }
  }
  if (lane == 0) { sval[warp] = val; sidx[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    val = lane < NT / 32 ? sval[lane] : -1.0;
    idx = lane < NT / 32 ? sidx[lane] : -1;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      double v2 = __shfl_down_sync(0xffffffffu, val, o);
      int i2 = __shfl_down_sync(0xffffffffu, idx, o);
      if (i2 >= 0 && (idx < 0 || v2 > val || (v2 == val && i2 < idx))) { val = v2; idx = i2; }
    }
    if (lane == 0) { sval[0] = val; sidx[0] = idx; }
  }
  __syncthreads();
  val = sval[0];
  idx = sidx[0];
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Panel factorisation (one CTA per front).
// ---------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) front_panel_kernel(const Front *__restrict__ fronts, int NB,
                                                         double pivtol) {
  const Front F = fronts[blockIdx.x];
  const int tid = threadIdx.x;
  const int n = F.n, nf = F.nf, ld = F.ld;
  double *__restrict__ A = F.A;
  double *__restrict__ W = F.W;
  const int k0 = F.state[ST_KCUR];
  if (k0 >= n) {
    if (tid == 0) F.state[ST_KPREV] = k0;
    return;
  }
  __shared__ double wrow[NBMAX];
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ double sh_d[4];

  const bool last_panel = (n - k0 <= NB);
  int k = k0;
  while (k < n && (last_panel || (k - k0) < NB - 1)) {
    const int kw = k - k0;
    double *__restrict__ Wk = W + (size_t)kw * ld;
    // ---- bring column k up to date:  W(:,kw) = A(:,k) - L(:,panel) * W(k,panel)^T ----
    for (int j = tid; j < kw; j += NT) wrow[j] = W[k + (size_t)j * ld];
    __syncthreads();
    double best = -1.0;
    int besti = -1;
    for (int i = k + tid; i < nf; i += NT) {
      double acc = A[i + (size_t)k * ld];
      const double *__restrict__ Li = A + i + (size_t)k0 * ld;
#pragma unroll 4
      for (int j = 0; j < kw; ++j) acc -= Li[(size_t)j * ld] * wrow[j];
      Wk[i] = acc;
      if (i > k && i < n) {
        const double a = fabs(acc);
        if (a > best) { best = a; besti = i; }
      }
      if (i == k) sh_d[0] = acc;
    }
    block_argmax<NT>(best, besti, sval, sidx);
    const double colmax = besti >= 0 ? best : 0.0;
    const int imax = besti;
    const double akk = sh_d[0];
    const double absakk = fabs(akk);

    int kstep = 1, kp = k;
    bool zero_pivot = false;
    if (!(fmax(absakk, colmax) > pivtol)) {
      zero_pivot = true;  // null (or NaN) column: record, keep going (LAPACK dsytf2 convention)
    } else if (absakk >= BK_ALPHA * colmax) {
      kp = k;
    } else {
      // ---- bring column imax up to date into W(:,kw+1) ----
      double *__restrict__ Wk1 = W + (size_t)(kw + 1) * ld;
      for (int j = tid; j < kw; j += NT) wrow[j] = W[imax + (size_t)j * ld];
      __syncthreads();
      double rbest = -1.0;
      int rbesti = -1;
      for (int i = k + tid; i < nf; i += NT) {
        double acc = (i < imax) ? A[imax + (size_t)i * ld] : A[i + (size_t)imax * ld];
        const double *__restrict__ Li = A + i + (size_t)k0 * ld;
#pragma unroll 4
        for (int j = 0; j < kw; ++j) acc -= Li[(size_t)j * ld] * wrow[j];
        Wk1[i] = acc;
        if (i < n && i != imax) {
          const double a = fabs(acc);
          if (a > rbest) { rbest = a; rbesti = i; }
        }
        if (i == imax) sh_d[1] = acc;
      }
      block_argmax<NT>(rbest, rbesti, sval, sidx);
      const double rowmax = rbesti >= 0 ? rbest : 0.0;
      const double wimax = sh_d[1];
      if (absakk >= BK_ALPHA * colmax * (colmax / rowmax)) {
        kp = k;
      } else if (fabs(wimax) >= BK_ALPHA * rowmax) {
        kp = imax;  // 1x1 pivot taken from the diagonal at imax
        for (int i = k + tid; i < nf; i += NT) Wk[i] = Wk1[i];
      } else {
        kp = imax;
        kstep = 2;
      }
    }
    __syncthreads();

    const int kk = k + kstep - 1;
    if (kp != kk) {
      // Symmetric interchange of rows/columns kk and kp in the not-yet-updated trailing matrix
      // (column kk itself is about to be overwritten by L), in the panel's L rows and in W.
      if (tid == 0) {
        A[kp + (size_t)kp * ld] = A[kk + (size_t)kk * ld];
        const int t = F.perm[kk];
        F.perm[kk] = F.perm[kp];
        F.perm[kp] = t;
      }
      for (int j = kk + 1 + tid; j < kp; j += NT) A[kp + (size_t)j * ld] = A[j + (size_t)kk * ld];
      for (int i = kp + 1 + tid; i < nf; i += NT) A[i + (size_t)kp * ld] = A[i + (size_t)kk * ld];
      for (int j = tid; j < kw; j += NT) {
        double *c = A + (size_t)(k0 + j) * ld;
        const double t = c[kk];
        c[kk] = c[kp];
        c[kp] = t;
      }
      for (int j = tid; j < kw + kstep; j += NT) {
        double *c = W + (size_t)j * ld;
        const double t = c[kk];
        c[kk] = c[kp];
        c[kp] = t;
      }
      __syncthreads();
    }

    if (kstep == 1) {
      const double d = Wk[k];
      const bool bad = zero_pivot || !(fabs(d) > pivtol) || !isfinite(d);
      const double rd = bad ? 0.0 : 1.0 / d;
      for (int i = k + 1 + tid; i < nf; i += NT) A[i + (size_t)k * ld] = Wk[i] * rd;
      if (tid == 0) {
        A[k + (size_t)k * ld] = bad ? 0.0 : d;
        F.ipiv[k] = kp;
        F.bsz[k] = 1;
        if (bad && F.state[ST_INFO] == 0) F.state[ST_INFO] = k + 1;
      }
    } else {
      const double *__restrict__ Wk1 = W + (size_t)(kw + 1) * ld;
      const double e11 = Wk[k], e21 = Wk[k + 1], e22 = Wk1[k + 1];
      // scaled inverse of the 2x2 pivot (as LAPACK dlasyf): robust when |e21| dominates
      const double d11 = e22 / e21, d22 = e11 / e21;
      const double t = 1.0 / (d11 * d22 - 1.0);
      const double s = t / e21;
      for (int i = k + 2 + tid; i < nf; i += NT) {
        const double w0 = Wk[i], w1 = Wk1[i];
        A[i + (size_t)k * ld] = s * (d11 * w0 - w1);
        A[i + (size_t)(k + 1) * ld] = s * (d22 * w1 - w0);
      }
      if (tid == 0) {
        A[k + (size_t)k * ld] = e11;
        A[k + 1 + (size_t)k * ld] = e21;
        A[k + 1 + (size_t)(k + 1) * ld] = e22;
        F.ipiv[k] = kp;
        F.ipiv[k + 1] = kp;
        F.bsz[k] = 2;
        F.bsz[k + 1] = 0;
        if (!isfinite(s) && F.state[ST_INFO] == 0) F.state[ST_INFO] = k + 1;
      }
    }
    __syncthreads();
    k += kstep;
  }
  if (tid == 0) {
    F.state[ST_KPREV] = k0;
    F.state[ST_KCUR] = k;
  }
}

// ---------------------------------------------------------------------------------------------
// Panel factorisation for tall fronts: one thread-block CLUSTER per front.  The rows of the panel are
// dealt over the CTAs of the cluster, so the bandwidth-bound column updates (each pivot step reads the
// panel computed so far) run on up to 8 SMs instead of one; arg-max results are exchanged through
// distributed shared memory, global data is handed over at cluster barriers.  Same algorithm and the
// same pivot choices as front_panel_kernel (ties resolve to the lowest row index).
// ---------------------------------------------------------------------------------------------
constexpr int PC_NT = 512;

// `extra` travels with the result: the value held by thread 0 of rank 0, or -- when s_extra is given -- the value some
// thread of rank `extra_rank` deposited in its CTA's shared slot *s_extra before the call.
__device__ __forceinline__ void cluster_argmax(cooperative_groups::cluster_group &cl, double &val, int &idx,
                                               double &extra, double *sval, int *sidx, double *xval, int *xidx,
                                               double *xextra, int &parity, const double *s_extra = nullptr,
                                               int extra_rank = 0) {
  // block-level reduce, then every CTA reads every CTA's result over DSMEM (one lane per remote CTA).
  // The exchange slots are double-buffered, so one cluster barrier per call suffices.
  block_argmax<PC_NT>(val, idx, sval, sidx);
  const int p = parity;
  parity ^= 1;
  if (threadIdx.x == 0) { xval[p] = val; xidx[p] = idx; xextra[p] = s_extra ? *s_extra : extra; }
  cl.sync();
  if (threadIdx.x < 32) {
    const unsigned nb = cl.num_blocks();
    double v = -1.0, e = 0.0;
    int i = -1;
    if (threadIdx.x < nb) {
      v = *cl.map_shared_rank(xval + p, threadIdx.x);
      i = *cl.map_shared_rank(xidx + p, threadIdx.x);
      if (threadIdx.x == 0) e = *cl.map_shared_rank(xextra + p, extra_rank);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double v2 = __shfl_down_sync(0xffffffffu, v, o);
      const int i2 = __shfl_down_sync(0xffffffffu, i, o);
      if (i2 >= 0 && (i < 0 || v2 > v || (v2 == v && i2 < i))) { v = v2; i = i2; }
    }
    if (threadIdx.x == 0) { sval[0] = v; sidx[0] = i; sval[1] = e; }
  }
  __syncthreads();
  val = sval[0];
  idx = sidx[0];
  extra = sval[1];
  __syncthreads();
}

__global__ void __launch_bounds__(PC_NT) front_panel_cluster_kernel(const Front *__restrict__ fronts, int NB,
                                                                    double pivtol) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const Front F = fronts[blockIdx.x / C];
  const int tid = threadIdx.x;
  const int gtid = rank * PC_NT + tid, GT = C * PC_NT;
  const int n = F.n, nf = F.nf, ld = F.ld;
  double *__restrict__ A = F.A;
  double *__restrict__ W = F.W;
  const int k0 = F.state[ST_KCUR];
  if (k0 >= n) {
    if (gtid == 0) F.state[ST_KPREV] = k0;
    return;  // uniform over the cluster
  }
  __shared__ double wrow[NBMAX];
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ double xval[2], xextra[2];
  __shared__ double s_wimax;
  __shared__ int xidx[2];
  int parity = 0;

  const bool last_panel = (n - k0 <= NB);
  int k = k0;
  while (k < n && (last_panel || (k - k0) < NB - 1)) {
    const int kw = k - k0;
    double *__restrict__ Wk = W + (size_t)kw * ld;
    for (int j = tid; j < kw; j += PC_NT) wrow[j] = W[k + (size_t)j * ld];
    __syncthreads();
    double best = -1.0, akk = 0.0;
    int besti = -1;
    for (int i = k + gtid; i < nf; i += GT) {
      double acc = A[i + (size_t)k * ld];
      const double *__restrict__ Li = A + i + (size_t)k0 * ld;
#pragma unroll 8
      for (int j = 0; j < kw; ++j) acc -= Li[(size_t)j * ld] * wrow[j];
      Wk[i] = acc;
      if (i > k && i < n) {
        const double a = fabs(acc);
        if (a > best) { best = a; besti = i; }
      }
      if (i == k) akk = acc;  // gtid 0 (rank 0, thread 0)
    }
    cluster_argmax(cl, best, besti, akk, sval, sidx, xval, xidx, xextra, parity);
    const double colmax = besti >= 0 ? best : 0.0;
    const int imax = besti;
    const double absakk = fabs(akk);

    int kstep = 1, kp = k;
    bool zero_pivot = false, copied = false;
    if (!(fmax(absakk, colmax) > pivtol)) {
      zero_pivot = true;
    } else if (absakk >= BK_ALPHA * colmax) {
      kp = k;
    } else {
      double *__restrict__ Wk1 = W + (size_t)(kw + 1) * ld;
      for (int j = tid; j < kw; j += PC_NT) wrow[j] = W[imax + (size_t)j * ld];
      __syncthreads();
      double rbest = -1.0, wimax = 0.0;
      int rbesti = -1;
      for (int i = k + gtid; i < nf; i += GT) {
        double acc = (i < imax) ? A[imax + (size_t)i * ld] : A[i + (size_t)imax * ld];
        const double *__restrict__ Li = A + i + (size_t)k0 * ld;
#pragma unroll 8
        for (int j = 0; j < kw; ++j) acc -= Li[(size_t)j * ld] * wrow[j];
        Wk1[i] = acc;
        if (i < n && i != imax) {
          const double a = fabs(acc);
          if (a > rbest) { rbest = a; rbesti = i; }
        }
        if (i == imax) s_wimax = acc;  // travels with the arg-max exchange: no CTA reads W(imax, kw+1) afterwards
      }
      cluster_argmax(cl, rbest, rbesti, wimax, sval, sidx, xval, xidx, xextra, parity, &s_wimax,
                     ((imax - k) % GT) / PC_NT);
      const double rowmax = rbesti >= 0 ? rbest : 0.0;
      if (absakk >= BK_ALPHA * colmax * (colmax / rowmax)) {
        kp = k;
      } else if (fabs(wimax) >= BK_ALPHA * rowmax) {
        kp = imax;
        for (int i = k + gtid; i < nf; i += GT) Wk[i] = Wk1[i];
        copied = true;
      } else {
        kp = imax;
        kstep = 2;
      }
    }
    // everything the interchange below reads was written before the barrier inside the last cluster_argmax --
    // except the copy above, whose rows belong to other CTAs (the decision is the same in every CTA)
    if (copied) cl.sync(); else __syncthreads();

    const int kk = k + kstep - 1;
    if (kp != kk) {
      if (gtid == 0) {
        A[kp + (size_t)kp * ld] = A[kk + (size_t)kk * ld];
        const int t = F.perm[kk];
        F.perm[kk] = F.perm[kp];
        F.perm[kp] = t;
      }
      for (int j = kk + 1 + gtid; j < kp; j += GT) A[kp + (size_t)j * ld] = A[j + (size_t)kk * ld];
      for (int i = kp + 1 + gtid; i < nf; i += GT) A[i + (size_t)kp * ld] = A[i + (size_t)kk * ld];
      for (int j = gtid; j < kw; j += GT) {
        double *c = A + (size_t)(k0 + j) * ld;
        const double t = c[kk];
        c[kk] = c[kp];
        c[kp] = t;
      }
      for (int j = gtid; j < kw + kstep; j += GT) {
        double *c = W + (size_t)j * ld;
        const double t = c[kk];
        c[kk] = c[kp];
        c[kp] = t;
      }
      cl.sync();
    }

    if (kstep == 1) {
      const double d = Wk[k];
      const bool bad = zero_pivot || !(fabs(d) > pivtol) || !isfinite(d);
      const double rd = bad ? 0.0 : 1.0 / d;
      for (int i = k + 1 + gtid; i < nf; i += GT) A[i + (size_t)k * ld] = Wk[i] * rd;
      if (gtid == 0) {
        A[k + (size_t)k * ld] = bad ? 0.0 : d;
        F.ipiv[k] = kp;
        F.bsz[k] = 1;
        if (bad && F.state[ST_INFO] == 0) F.state[ST_INFO] = k + 1;
      }
    } else {
      const double *__restrict__ Wk1 = W + (size_t)(kw + 1) * ld;
      const double e11 = Wk[k], e21 = Wk[k + 1], e22 = Wk1[k + 1];
      const double d11 = e22 / e21, d22 = e11 / e21;
      const double t = 1.0 / (d11 * d22 - 1.0);
      const double sc = t / e21;
      for (int i = k + 2 + gtid; i < nf; i += GT) {
        const double w0 = Wk[i], w1 = Wk1[i];
        A[i + (size_t)k * ld] = sc * (d11 * w0 - w1);
        A[i + (size_t)(k + 1) * ld] = sc * (d22 * w1 - w0);
      }
      if (gtid == 0) {
        A[k + (size_t)k * ld] = e11;
        A[k + 1 + (size_t)k * ld] = e21;
        A[k + 1 + (size_t)(k + 1) * ld] = e22;
        F.ipiv[k] = kp;
        F.ipiv[k + 1] = kp;
        F.bsz[k] = 2;
        F.bsz[k + 1] = 0;
        if (!isfinite(sc) && F.state[ST_INFO] == 0) F.state[ST_INFO] = k + 1;
      }
    }
    // No cluster barrier at the end of a column: the L entries a thread reads in the next sweep are the ones it just
    // wrote (same row mapping), W(:, kw) and W(:, kw + 1) were published by the barriers inside cluster_argmax, the
    // candidate's diagonal entry travelled with the exchange, and a copy or an interchange is followed by its own
    // cluster barrier above.
    __syncthreads();
    k += kstep;
  }
  if (gtid == 0) {
    F.state[ST_KPREV] = k0;
    F.state[ST_KCUR] = k;
  }
  if (C > 1) cl.sync();   // see front_panel_cluster_oc_kernel: nobody leaves while a peer may still read its candidates
}

// Apply the interchanges of the last panel to the columns of L left of it (one thread per column).
__global__ void front_swaps_left_kernel(const Front *__restrict__ fronts) {
  const Front F = fronts[blockIdx.y];
  const int k0 = F.state[ST_KPREV], k1 = F.state[ST_KCUR];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k0 || k0 == k1) return;
  double *col = F.A + (size_t)c * F.ld;
  for (int k = k0; k < k1; ++k) {
    const int b = F.bsz[k];
    if (b == 0) continue;
    const int kk = k + b - 1, kp = F.ipiv[k];
    if (kp != kk) {
      const double t = col[kk];
      col[kk] = col[kp];
      col[kp] = t;
    }
  }
}

// The same panel factorisation with the panel's rows of L kept on chip (see inside).  Used when fronts are tall and few.
#ifdef PP_TRACE_SOLVE
__device__ long long g_pc_trace[24];
#define PC_TR(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t_ = clock64(); g_pc_trace[k] += t_ - t_last; t_last = t_; } } while (0)
#define PC_CNT(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_pc_trace[k] += 1; } while (0)
#else
#define PC_TR(k) do { } while (0)
#define PC_CNT(k) do { } while (0)
#endif
// exact maximum over a warp of the bit patterns of non-negative doubles (two 32-bit redux.sync operations)
__device__ __forceinline__ unsigned long long warp_max_bits(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned lo = hi == mh ? (unsigned)v : 0u;
  const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
  return ((unsigned long long)mh << 32) | ml;
}

constexpr int OC_REG = 16, OC_SM = NBMAX - OC_REG;
constexpr size_t OC_SMEM = ((size_t)OC_SM * PC_NT + (size_t)NBMAX * NBMAX) * sizeof(double);

__global__ void __launch_bounds__(PC_NT) front_panel_cluster_oc_kernel(const Front *__restrict__ fronts, int NB,
                                                                    double pivtol, int spec) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const Front F = fronts[blockIdx.x / C];
  const int tid = threadIdx.x;
  const int gtid = rank * PC_NT + tid, GT = C * PC_NT;
  const int n = F.n, nf = F.nf, ld = F.ld;
  double *__restrict__ A = F.A;
  double *__restrict__ W = F.W;
  const int k0 = F.state[ST_KCUR];
  if (k0 >= n) {
    if (gtid == 0) F.state[ST_KPREV] = k0;
    return;  // uniform over the cluster
  }
  __shared__ double wrow[NBMAX];
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ double xval[2], xextra[2];
  __shared__ double s_wimax, s_akk;
  __shared__ int xidx[2];
  int parity = 0;
  // The thread's row of the current panel stays on chip: columns [0, OC_REG) in registers, the rest in shared memory
  // ([column][thread]: conflict-free).  Rows are dealt once per panel (row0 = k0 + gtid), not per column, so the
  // lazy sweep of a column is kw fused multiply-adds from registers / shared memory instead of kw loads from L2
  // (130 kB per CTA and column at 4 000 rows).  Rows beyond the first GT of a very tall front use the L2 path.
  extern __shared__ double Ls[];  // [OC_SM][PC_NT], then (speculative prologue) the NBMAX x NBMAX diagonal block
  double *const Wd = Ls + (size_t)OC_SM * PC_NT;
  __shared__ unsigned long long s_colmax[NBMAX];
  __shared__ double s_rd[NBMAX];
  __shared__ unsigned s_bad[2];
  __shared__ int s_nacc;
  double Lr[OC_REG];
#pragma unroll
  for (int j = 0; j < OC_REG; ++j) Lr[j] = 0.0;
  const int row0 = k0 + gtid;
  auto dot_row = [&](int kw, const double *__restrict__ wr) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < OC_REG; ++j)
      if (j < kw) acc += Lr[j] * wr[j];
    for (int j = OC_REG; j < kw; ++j) acc += Ls[(j - OC_REG) * PC_NT + tid] * wr[j];
    return acc;
  };
  auto dot_onchip = [&](int kw) { return dot_row(kw, wrow); };
  auto store_onchip = [&](int slot, double v) {
    if (slot < OC_REG) {
#pragma unroll
      for (int j = 0; j < OC_REG; ++j)
        if (j == slot) Lr[j] = v;
    } else {
      Ls[(slot - OC_REG) * PC_NT + tid] = v;
    }
  };

  const bool last_panel = (n - k0 <= NB);
  int k = k0;
  // Speculation pays when the pivots are benign (the generator's KKT blocks: every diagonal passes); on fronts whose
  // tests fail early (late-IPM blocks with delayed pivots) the attempt is wasted, so a panel that committed less than
  // half of its columns switches speculation off for the front's next three panels (state[3], reset per factorisation).
  const int spec_penalty = F.state[3];
  int spec_next = spec_penalty > 0 ? spec_penalty - 1 : 0;
  if (spec && spec_penalty == 0) {
    // ---- speculative panel ---------------------------------------------------------------------------------
    // Measured on the 4 082 columns of a config-5 root and its coupling front: every column accepts its diagonal at
    // once (no second column, no interchange, no 2x2 pivot), and each still pays a cluster-wide arg-max.  So the
    // panel is first factorised AS IF every diagonal passed: (A) the CTA that owns the panel's rows factorises the
    // w x w diagonal block in shared memory (right-looking, one barrier per column); (B) every row then runs its own
    // w-column recurrence
    //   W(i,c) = A(i,c) - sum_{j<c} L(i,j) W(c,j),  L(i,c) = W(i,c) / d_c
    // against that block with no communication at all, row chunk by row chunk with the chunk's L on chip, writing W
    // and folding |W(i,c)| into per-column maxima; (C) ONE reduction over the cluster gives all w column maxima and
    // the Bunch-Kaufman tests are checked in order.  Columns before the first failure are exactly what the
    // column-by-column loop below would have produced (same tests, same pivots) and are committed; from the failing
    // column on that loop takes over for the rest of the panel.
    const int w = last_panel ? n - k0 : NB - 1;
#ifdef PP_TRACE_SOLVE
    long long t_last = clock64();
#endif
    PC_CNT(23);
    for (int idx = tid; idx < NBMAX * NBMAX; idx += PC_NT) {
      const int r = idx % NBMAX, c = idx / NBMAX;   // Wd[c * NBMAX + r] = entry (row k0 + r, column k0 + c), r >= c
      Wd[idx] = (r < w && c <= r) ? A[(k0 + r) + (size_t)(k0 + c) * ld] : 0.0;
    }
    if (tid < NBMAX) s_colmax[tid] = 0ull;
    __syncthreads();
    PC_TR(12);
    if (rank == 0) {
      // (A) the diagonal block, right-looking in shared memory with the whole CTA: at step c the remaining lower
      // triangle takes the rank-one update of column c (every entry receives its updates in ascending column order,
      // as the rows below do), one barrier per column; column c is left holding W(:, c)
      for (int c = 0; c < w; ++c) {
        const double rd = __drcp_rn(Wd[c * NBMAX + c]);
        if (tid == 0) s_rd[c] = rd;
        const int m = w - 1 - c;
        for (int idx = tid; idx < m * m; idx += PC_NT) {
          const int rr = c + 1 + idx % m, cc = c + 1 + idx / m;   // consecutive threads walk down a column
          if (cc <= rr) Wd[cc * NBMAX + rr] -= (Wd[c * NBMAX + rr] * rd) * Wd[c * NBMAX + cc];
        }
        __syncthreads();
      }
      // column maxima of the block's part of the tests, and the block itself (W of the panel's rows) to global
      // memory: for the other CTAs and for the update kernel
      for (int c = tid >> 5; c < w; c += PC_NT / 32) {
        const int l = tid & 31;
        double mx = 0.0;
        for (int r = c + 1 + l; r < w && k0 + r < n; r += 32) mx = fmax(mx, fabs(Wd[c * NBMAX + r]));
        const unsigned long long mb = warp_max_bits((unsigned long long)__double_as_longlong(mx));
        if (l == 0) s_colmax[c] = mb;
      }
      for (int idx = tid; idx < NBMAX * NBMAX; idx += PC_NT) {
        const int r = idx % NBMAX, c = idx / NBMAX;
        if (r < w && c <= r) W[(k0 + r) + (size_t)c * ld] = Wd[idx];
      }
    }
    PC_TR(13);
    if (C > 1) { __threadfence(); cl.sync(); } else __syncthreads();
    PC_TR(14);
    if (rank != 0) {
      for (int idx = tid; idx < NBMAX * NBMAX; idx += PC_NT) {
        const int r = idx % NBMAX, c = idx / NBMAX;
        Wd[idx] = (r < w && c <= r) ? W[(k0 + r) + (size_t)c * ld] : 0.0;
      }
      __syncthreads();
    }
    if (rank != 0 && tid < NBMAX) s_rd[tid] = tid < w ? __drcp_rn(Wd[tid * NBMAX + tid]) : 0.0;
    __syncthreads();
    // (B) rows below the block, one chunk of GT rows at a time (the chunk's L lives in Lr / Ls); rows are dealt as
    // in the loop below: row0, row0 + GT, ...
    for (int i = row0;; i += GT) {
      if (!__syncthreads_or(i < nf)) break;   // uniform exit: every thread of the CTA is past the last row
      const bool on = i >= k0 + w && i < nf;
      // eight columns at a time: the part against the columns before the block is a small GEMV in which one entry
      // of this row's L meets eight multipliers (two 16-byte broadcast loads per pair), then the 8 x 8 triangle
      for (int c0 = 0; c0 < w; c0 += 8) {
        double acc[8], lb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = (on && c0 + q < w) ? A[i + (size_t)(k0 + c0 + q) * ld] : 0.0;
        auto gemv8 = [&](double lj, int j) {
          const double2 *mp = reinterpret_cast<const double2 *>(Wd + j * NBMAX + c0);   // W(c0 .. c0+7, j)
          const double2 m0 = mp[0], m1 = mp[1], m2 = mp[2], m3 = mp[3];
          acc[0] -= lj * m0.x; acc[1] -= lj * m0.y; acc[2] -= lj * m1.x; acc[3] -= lj * m1.y;
          acc[4] -= lj * m2.x; acc[5] -= lj * m2.y; acc[6] -= lj * m3.x; acc[7] -= lj * m3.y;
        };
        if (c0 >= 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gemv8(Lr[j], j);
        }
        if (c0 >= 16) {
#pragma unroll
          for (int j = 8; j < 16; ++j) gemv8(Lr[j], j);
        }
        for (int j = OC_REG; j < c0; ++j) gemv8(Ls[(j - OC_REG) * PC_NT + tid], j);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = c0 + q;
#pragma unroll
          for (int pq = 0; pq < 8; ++pq)
            if (pq < q) acc[q] -= lb[pq] * Wd[(c0 + pq) * NBMAX + c];
          lb[q] = c < w ? acc[q] * s_rd[c] : 0.0;
          if (on && c < w) {
            W[i + (size_t)c * ld] = acc[q];
            store_onchip(c, lb[q]);
          }
          const unsigned long long mb =
              warp_max_bits((on && c < w && i < n) ? (unsigned long long)__double_as_longlong(fabs(acc[q])) : 0ull);
          if ((tid & 31) == 0 && mb > *((volatile unsigned long long *)&s_colmax[c < w ? c : 0])) atomicMax(&s_colmax[c], mb);
        }
      }
    }
    __syncthreads();
    PC_TR(15);
    // (C) column maxima over the cluster (the chunk scratch is free now: Ls doubles as the exchange area)
    unsigned long long *xcm = reinterpret_cast<unsigned long long *>(Ls);
    if (C > 1) {
      cl.sync();   // every CTA is done with its Ls before anybody writes into it remotely
      if (tid < NBMAX)
        for (int q = 0; q < C; ++q) cl.map_shared_rank(xcm, q)[rank * NBMAX + tid] = s_colmax[tid];
      cl.sync();
      if (tid < NBMAX) {
        unsigned long long m = 0ull;
        for (int q = 0; q < C; ++q) m = max(m, xcm[q * NBMAX + tid]);
        s_colmax[tid] = m;
      }
      __syncthreads();
    }
    if (tid < NBMAX) {
      bool ok = true;
      if (tid < w) {
        const double d = Wd[tid * NBMAX + tid], absakk = fabs(d);
        const double colmax = __longlong_as_double((long long)s_colmax[tid]);
        ok = (fmax(absakk, colmax) > pivtol) && (absakk >= BK_ALPHA * colmax) && (absakk > pivtol) && isfinite(d);
      }
      const unsigned bad = __ballot_sync(0xffffffffu, !ok);
      if ((tid & 31) == 0) s_bad[tid >> 5] = bad;
    }
    __syncthreads();
    if (tid == 0) s_nacc = s_bad[0] ? __ffs(s_bad[0]) - 1 : (s_bad[1] ? 32 + __ffs(s_bad[1]) - 1 : w);
    __syncthreads();
    const int nacc = s_nacc;
    if (2 * nacc < w) spec_next = 3;
    PC_TR(16);
    // commit the accepted columns: L = W / d (the same expression the column-by-column loop uses), pivot records
    for (int i = k0 + 1 + gtid; i < nf; i += GT) {   // row i gets its entries of the columns c < min(nacc, i - k0)
      const int cend = min(nacc, i - k0);
      for (int c0 = 0; c0 < cend; c0 += 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = c0 + q < cend ? W[i + (size_t)(c0 + q) * ld] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (c0 + q < cend) A[i + (size_t)(k0 + c0 + q) * ld] = v[q] * s_rd[c0 + q];
      }
    }
    if (gtid < nacc) {
      A[(k0 + gtid) + (size_t)(k0 + gtid) * ld] = Wd[gtid * NBMAX + gtid];
      F.ipiv[k0 + gtid] = k0 + gtid;
      F.bsz[k0 + gtid] = 1;
    }
    k = k0 + nacc;
    PC_TR(17);
    if (k < n && (last_panel || (k - k0) < NB - 1)) {
      // a column failed its test: the general loop continues from it; this thread's first row goes back on chip
      if (C > 1) { __threadfence(); cl.sync(); } else __syncthreads();
#pragma unroll
      for (int j = 0; j < OC_REG; ++j) Lr[j] = 0.0;
      if (row0 < nf)
        for (int j = 0; j < nacc; ++j) store_onchip(j, row0 > k0 + j ? A[row0 + (size_t)(k0 + j) * ld] : 0.0);
      __syncthreads();
    }
  }
#ifdef PP_TRACE_SOLVE
  long long t_last = clock64();
#endif
  while (k < n && (last_panel || (k - k0) < NB - 1)) {
    const int kw = k - k0;
    double *__restrict__ Wk = W + (size_t)kw * ld;
    PC_TR(0);
    PC_CNT(8);
    for (int j = tid; j < kw; j += PC_NT) wrow[j] = W[k + (size_t)j * ld];
    __syncthreads();
    PC_TR(1);
    double best = -1.0, akk = 0.0;
    int besti = -1;
    for (int i = row0; i < nf; i += GT) {
      if (i < k) continue;  // a row of this panel that is already a pivot row
      double acc = A[i + (size_t)k * ld];
      if (i == row0) {
        acc -= dot_onchip(kw);
      } else {
        const double *__restrict__ Li = A + i + (size_t)k0 * ld;
#pragma unroll 8
        for (int j = 0; j < kw; ++j) acc -= Li[(size_t)j * ld] * wrow[j];
      }
      Wk[i] = acc;
      if (i > k && i < n) {
        const double a = fabs(acc);
        if (a > best) { best = a; besti = i; }
      }
      if (i == k) s_akk = acc;  // travels with the exchange from the CTA that owns row k
    }
    PC_TR(2);
    cluster_argmax(cl, best, besti, akk, sval, sidx, xval, xidx, xextra, parity, &s_akk, ((k - k0) % GT) / PC_NT);
    PC_TR(3);
    const double colmax = besti >= 0 ? best : 0.0;
    const int imax = besti;
    const double absakk = fabs(akk);

    int kstep = 1, kp = k;
    bool zero_pivot = false, copied = false;
    if (!(fmax(absakk, colmax) > pivtol)) {
      zero_pivot = true;
    } else if (absakk >= BK_ALPHA * colmax) {
      kp = k;
    } else {
      double *__restrict__ Wk1 = W + (size_t)(kw + 1) * ld;
      PC_CNT(9);
      for (int j = tid; j < kw; j += PC_NT) wrow[j] = W[imax + (size_t)j * ld];
      __syncthreads();
      double rbest = -1.0, wimax = 0.0;
      int rbesti = -1;
      for (int i = row0; i < nf; i += GT) {
        if (i < k) continue;
        double acc = (i < imax) ? A[imax + (size_t)i * ld] : A[i + (size_t)imax * ld];
        if (i == row0) {
          acc -= dot_onchip(kw);
        } else {
          const double *__restrict__ Li = A + i + (size_t)k0 * ld;
#pragma unroll 8
          for (int j = 0; j < kw; ++j) acc -= Li[(size_t)j * ld] * wrow[j];
        }
        Wk1[i] = acc;
        if (i < n && i != imax) {
          const double a = fabs(acc);
          if (a > rbest) { rbest = a; rbesti = i; }
        }
        if (i == imax) s_wimax = acc;  // travels with the arg-max exchange: no CTA reads W(imax, kw+1) afterwards
      }
      cluster_argmax(cl, rbest, rbesti, wimax, sval, sidx, xval, xidx, xextra, parity, &s_wimax,
                     ((imax - k0) % GT) / PC_NT);
      const double rowmax = rbesti >= 0 ? rbest : 0.0;
      if (absakk >= BK_ALPHA * colmax * (colmax / rowmax)) {
        kp = k;
      } else if (fabs(wimax) >= BK_ALPHA * rowmax) {
        kp = imax;
        for (int i = row0; i < nf; i += GT)
          if (i >= k) Wk[i] = Wk1[i];
        copied = true;
      } else {
        kp = imax;
        kstep = 2;
      }
    }
    // everything the interchange below reads was written before the barrier inside the last cluster_argmax --
    // except the copy above, whose rows belong to other CTAs (the decision is the same in every CTA)
    PC_TR(4);
    if (copied) cl.sync(); else __syncthreads();

    const int kk = k + kstep - 1;
    if (kp != kk) {
      PC_CNT(10);
      if (gtid == 0) {
        A[kp + (size_t)kp * ld] = A[kk + (size_t)kk * ld];
        const int t = F.perm[kk];
        F.perm[kk] = F.perm[kp];
        F.perm[kp] = t;
      }
      for (int j = kk + 1 + gtid; j < kp; j += GT) A[kp + (size_t)j * ld] = A[j + (size_t)kk * ld];
      for (int i = kp + 1 + gtid; i < nf; i += GT) A[i + (size_t)kp * ld] = A[i + (size_t)kk * ld];
      for (int j = gtid; j < kw; j += GT) {
        double *c = A + (size_t)(k0 + j) * ld;
        const double t = c[kk];
        c[kk] = c[kp];
        c[kp] = t;
      }
      for (int j = gtid; j < kw + kstep; j += GT) {
        double *c = W + (size_t)j * ld;
        const double t = c[kk];
        c[kk] = c[kp];
        c[kp] = t;
      }
      cl.sync();
      // the two exchanged rows changed owners' contents: their on-chip copies are read back
      if (row0 == kk || row0 == kp)
        for (int j = 0; j < kw; ++j) store_onchip(j, A[row0 + (size_t)(k0 + j) * ld]);
    }

    PC_TR(5);
    if (kstep == 2) PC_CNT(11);
    if (kstep == 1) {
      const double d = Wk[k];
      const bool bad = zero_pivot || !(fabs(d) > pivtol) || !isfinite(d);
      const double rd = bad ? 0.0 : 1.0 / d;
      for (int i = row0; i < nf; i += GT) {
        if (i <= k) continue;
        const double v = Wk[i] * rd;
        A[i + (size_t)k * ld] = v;
        if (i == row0) store_onchip(kw, v);
      }
      if (gtid == 0) {
        A[k + (size_t)k * ld] = bad ? 0.0 : d;
        F.ipiv[k] = kp;
        F.bsz[k] = 1;
        if (bad && F.state[ST_INFO] == 0) F.state[ST_INFO] = k + 1;
      }
    } else {
      const double *__restrict__ Wk1 = W + (size_t)(kw + 1) * ld;
      const double e11 = Wk[k], e21 = Wk[k + 1], e22 = Wk1[k + 1];
      const double d11 = e22 / e21, d22 = e11 / e21;
      const double t = 1.0 / (d11 * d22 - 1.0);
      const double sc = t / e21;
      for (int i = row0; i < nf; i += GT) {
        if (i <= k + 1) continue;
        const double w0 = Wk[i], w1 = Wk1[i];
        const double l0 = sc * (d11 * w0 - w1), l1 = sc * (d22 * w1 - w0);
        A[i + (size_t)k * ld] = l0;
        A[i + (size_t)(k + 1) * ld] = l1;
        if (i == row0) { store_onchip(kw, l0); store_onchip(kw + 1, l1); }
      }
      if (gtid == 0) {
        A[k + (size_t)k * ld] = e11;
        A[k + 1 + (size_t)k * ld] = e21;
        A[k + 1 + (size_t)(k + 1) * ld] = e22;
        F.ipiv[k] = kp;
        F.ipiv[k + 1] = kp;
        F.bsz[k] = 2;
        F.bsz[k + 1] = 0;
        if (!isfinite(sc) && F.state[ST_INFO] == 0) F.state[ST_INFO] = k + 1;
      }
    }
    // No cluster barrier at the end of a column: the L entries a thread reads in the next sweep are its own (rows are
    // dealt once per panel), W(:, kw) and W(:, kw + 1) were published by the barriers inside cluster_argmax, the
    // candidate's diagonal entry travelled with the exchange, and a copy or an interchange is followed by its own
    // cluster barrier above.
    __syncthreads();
    PC_TR(6);
    k += kstep;
  }
  if (gtid == 0) {
    F.state[ST_KPREV] = k0;
    F.state[ST_KCUR] = k;
    F.state[3] = spec_next;
  }
  // cluster_argmax PULLS the candidates out of the peers' shared memory after its barrier: a CTA that ran ahead must
  // not leave while a peer may still be reading (distributed shared memory lives only as long as its CTA does)
  if (C > 1) cl.sync();
}

// ---------------------------------------------------------------------------------------------
// Trailing update  A(i,j) -= sum_k L(i,k) * W(j,k)   (i >= j >= kcur), FP64 DMMA m8n8k4.
// 128 x 128 output tile per CTA, 16 warps of 32 x 32, whole panel (K <= NBMAX) staged in smem.
// ---------------------------------------------------------------------------------------------
constexpr int UT = 128;          // tile rows
constexpr int UTN = 64;          // tile columns
constexpr int UROW = UT + 4;     // smem pitches (doubles): pitch % 16 == 4 -> conflict-free fragment loads
constexpr int UCOL = UTN + 4;
constexpr int UPD_THREADS = 256;
constexpr size_t UPD_SMEM = (size_t)NBMAX * (UROW + UCOL) * sizeof(double);  // 100 KB: two CTAs per SM

#ifdef PP_UPDATE_PROBE
__device__ int g_update_probe = 0;
#define UPD_PROBE(bit) (g_update_probe & (bit))
#else
#define UPD_PROBE(bit) 0
#endif

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

// number of (128 x 64) tiles of the lower-triangular trailing region that starts at `kcur` in a
// front of `nf` rows; tiles are anchored at absolute multiples of the tile size
__host__ __device__ inline int update_tile_count(int nf, int kcur) {
  const int t0 = kcur / UT, c0 = kcur / UTN, tend = (nf + UT - 1) / UT;
  int total = 0;
  for (int t = t0; t < tend; ++t) total += min(2 * t + 2, (nf + UTN - 1) / UTN) - c0;
  return total;
}

__global__ void __launch_bounds__(UPD_THREADS, 2) front_update_kernel(const Front *__restrict__ fronts) {
  const Front F = fronts[blockIdx.y];
  const int kprev = F.state[ST_KPREV], kcur = F.state[ST_KCUR];
  const int kw = kcur - kprev;
  const int nf = F.nf, ld = F.ld;
  if (kw == 0 || kcur >= nf) return;
  // linear index -> (row tile ti, column tile tj) with 64*tj <= 128*ti + 127
  const int t0 = kcur / UT, c0 = kcur / UTN, tend = (nf + UT - 1) / UT, cend = (nf + UTN - 1) / UTN;
  int rem = blockIdx.x, ti = t0;
  for (; ti < tend; ++ti) {
    const int cnt = min(2 * ti + 2, cend) - c0;
    if (rem < cnt) break;
    rem -= cnt;
  }
  if (ti >= tend) return;
  const int tj = c0 + rem;
  const int row0 = ti * UT, col0 = tj * UTN;
  // rows / columns [n, nb) are the unused delayed-pivot slots: exactly zero in A, L and W, nothing to update
  if ((row0 >= F.n && row0 + UT <= F.nb) || (col0 >= F.n && col0 + UTN <= F.nb)) return;

  extern __shared__ __align__(16) double smem[];
  double *sA = smem;                 // [kpad][UROW]  L(row0 + r, kprev + k)
  double *sB = smem + NBMAX * UROW;  // [kpad][UCOL]  W(col0 + c, k)
  const int kpad = (kw + 3) & ~3;
  const double *gA = F.A + (size_t)kprev * ld + row0;
  const double *gB = F.W + col0;
  for (int idx = threadIdx.x; idx < kpad * (UT / 2); idx += UPD_THREADS) {
    const int k = idx / (UT / 2), r = (idx % (UT / 2)) * 2;
    const bool ok = k < kw && row0 + r < nf;
    cp_async16(sA + k * UROW + r, ok ? gA + (size_t)k * ld + r : F.A, ok);
  }
  for (int idx = threadIdx.x; idx < kpad * (UTN / 2); idx += UPD_THREADS) {
    const int k = idx / (UTN / 2), r = (idx % (UTN / 2)) * 2;
    const bool ok = k < kw && col0 + r < nf;
    cp_async16(sB + k * UCOL + r, ok ? gB + (size_t)k * ld + r : F.A, ok);
  }
  asm volatile("cp.async.commit_group;\n" ::);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps of 32 x 32
  const bool active = row0 + wm * 32 + 31 >= col0 + wn * 32;  // not strictly above the diagonal
  const int g = lane >> 2, q = lane & 3;
  const int rbase = row0 + wm * 32 + g;
  const int cbase = col0 + wn * 32 + q * 2;
  double acc[4][4][2];
  if (active) {
#pragma unroll
    for (int fm = 0; fm < 4; ++fm)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int r = rbase + fm * 8, c = cbase + fn * 8 + e;
          const bool ok = r < nf && c >= kcur && r >= c && !UPD_PROBE(1);
          acc[fm][fn][e] = ok ? F.A[r + (size_t)c * ld] : 0.0;
        }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  if (!active) return;

  const double *pa = sA + q * UROW + wm * 32 + g;
  const double *pb = sB + q * UCOL + wn * 32 + g;
  for (int k = 0; k < (UPD_PROBE(8) ? 4 : kpad); k += 4) {
    double a[4], b[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      a[f] = -pa[k * UROW + f * 8];
      b[f] = pb[k * UCOL + f * 8];
    }
#pragma unroll
    for (int fm = 0; fm < 4; ++fm)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a[fm], b[fn]);
  }
#pragma unroll
  for (int fm = 0; fm < 4; ++fm)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = rbase + fm * 8, c = cbase + fn * 8 + e;
        if (r < nf && c >= kcur && r >= c && !(UPD_PROBE(2) && acc[fm][fn][e] != 1.2345e300)) F.A[r + (size_t)c * ld] = acc[fm][fn][e];
      }
}

// ---------------------------------------------------------------------------------------------
// The same update, software-pipelined.  One CTA per *strip*: up to `S` consecutive 128 x 64 tiles of one tile row.
//
// What bounds the one-tile kernel above (tools/update_probe.cu, 32 fronts of the config-5 root shape, one launch):
// full 1.06 ms; staging of the L / W panels alone 0.35 ms (96 KB per tile out of L2 = 4.1 TB/s); C in + out without
// the k-loop 0.78 ms; k-loop + staging without any C traffic 0.89 ms; the useful flops at the DMMA peak 0.52 ms.
// Every tile moves 224 KB between L2 and the SM for 1.05 MFLOP, nothing of it under the tile's own k-loop.
// Here: the strip's 128 rows of L (A operand, 64 KB) are staged once and stay in shared memory (5 KB per tile at
// S = 12 instead of 64); the W rows (B operand, 32 KB) and the OLD VALUES OF THE C TILE (64 KB) of tile t + 1 arrive by
// cp.async -- 16 B per thread, whole 128 B lines per request, two running pointers per thread and tile -- while the
// tensor pipe works on tile t, so a tile costs 165 KB of traffic.  At the start of a tile the accumulators are loaded
// from the staged C tile (acc = C, acc += (-L) W^T with the staged L negated once per strip: the summation order of the
// kernel above, results are bit-identical), the k-loop reads double-buffered fragments from shared memory, and the
// results of an interior tile leave from registers four at a time between the DMMA batches of the next tile's k-loop.
// 204 KB of shared memory, one CTA per SM, 8 warps of 32 x 32.
// Measured (same probe): 0.96 ms at S = 16, 20.0 TFLOP/s of useful flops = 56 % of cuBLAS DGEMM (22 TFLOP/s counting the
// diagonal / edge tiles' masked work); ncu stall samples: k-loop 62 %, start of a strip (A panel, nothing to overlap
// it with at one CTA per SM) 11 %, per-tile staging issue / barriers / accumulator loads 22 %.  A rank-64 update is
// balanced between the tensor pipe (0.57 ms) and the C stream (2.6 GB, 0.57 ms at 4.5 TB/s): what is left needs fewer
// passes over C (rank-128 updates), not a better pass.
// ---------------------------------------------------------------------------------------------
constexpr int UCP = UT + 2;       // pitch of the staged C tile: 2 * UCP % 16 == 4 -> conflict-free accumulator loads
constexpr size_t UPS_SMEM = ((size_t)NBMAX * (UROW + 2 * UCOL) + (size_t)UTN * UCP) * sizeof(double);  // 203 776 B
constexpr int UPS_MAX = 16;  // longest strip (tiles)

// strips of the trailing region that starts at `kcur`: every tile row's tiles are cut into ceil(cnt / S) nearly equal runs
__host__ __device__ inline int update_strip_count(int nf, int kcur, int S) {
  const int t0 = kcur / UT, c0 = kcur / UTN, tend = (nf + UT - 1) / UT, cend = (nf + UTN - 1) / UTN;
  int total = 0;
  for (int t = t0; t < tend; ++t) total += (min(2 * t + 2, cend) - c0 + S - 1) / S;
  return total;
}

__global__ void __launch_bounds__(UPD_THREADS, 1) front_update_strip_kernel(const Front *__restrict__ fronts, int S) {
  const Front F = fronts[blockIdx.y];
  const int kprev = F.state[ST_KPREV], kcur = F.state[ST_KCUR];
  const int kw = kcur - kprev;
  const int nf = F.nf, ld = F.ld;
  if (kw == 0 || kcur >= nf) return;
  const int t0 = kcur / UT, c0 = kcur / UTN, tend = (nf + UT - 1) / UT, cend = (nf + UTN - 1) / UTN;
  int rem = blockIdx.x, ti = t0, cnt = 0, ns = 0;
  for (; ti < tend; ++ti) {
    cnt = min(2 * ti + 2, cend) - c0;
    ns = (cnt + S - 1) / S;
    if (rem < ns) break;
    rem -= ns;
  }
  if (ti >= tend) return;
  // run `rem` of `ns` nearly equal runs of the row's `cnt` tiles
  const int base = cnt / ns, extra = cnt % ns;
  const int tj0 = c0 + rem * base + min(rem, extra);
  const int ntj = base + (rem < extra ? 1 : 0);
  const int row0 = ti * UT;
  // rows / columns [n, nb) are the unused delayed-pivot slots: exactly zero in A, L and W, nothing to update
  const int gap0 = F.n, gap1 = F.nb;
  if (row0 >= gap0 && row0 + UT <= gap1) return;

  extern __shared__ __align__(16) double smem[];
  double *sA = smem;                                  // [kpad][UROW]     -L(row0 + r, kprev + k)
  double *sB = smem + NBMAX * UROW;                   // 2 x [kpad][UCOL] W(col0 + c, k) of the current / next tile
  double *sC = smem + NBMAX * (UROW + 2 * UCOL);      // [UTN][UCP]       old A(row0 + r, col0 + c) of the NEXT tile
  const int kpad = (kw + 7) & ~7;
  const int tid = threadIdx.x;
  {
    const double *gA = F.A + (size_t)kprev * ld + row0;
    for (int idx = tid; idx < kpad * (UT / 2); idx += UPD_THREADS) {
      const int k = idx / (UT / 2), r = (idx % (UT / 2)) * 2;
      const bool ok = k < kw && row0 + r < nf;
      cp_async16(sA + k * UROW + r, ok ? gA + (size_t)k * ld + r : F.A, ok);
    }
  }
  auto skipped = [&](int tj) { return tj * UTN >= gap0 && tj * UTN + UTN <= gap1; };
  // Staging of tile tj: W rows and old C values, one cp.async group.  Thread `tid` always moves the same 16-byte
  // column of the W block (rows k = tid / 32 + 8 j) and of the C tile (columns c = tid / 64 + 4 j): two running
  // pointers per tile, no index arithmetic per request.
  const int brow = tid >> 5, bcol = (tid & 31) * 2;   // W: k = brow + 8 j, entries col0 + bcol, + 1
  const int ccol = tid >> 6, crow = (tid & 63) * 2;   // C: column col0 + ccol + 4 j, rows row0 + crow, + 1
  const double *const wsrc0 = F.W + (size_t)brow * ld + bcol;
  const double *const csrc0 = F.A + (size_t)ccol * ld + row0 + crow;
  const bool full_k = kw == kpad;
  auto stage = [&](int tj, int buf) {
    if (!skipped(tj)) {
      const int col0 = tj * UTN;
      double *dst = sB + buf * NBMAX * UCOL + brow * UCOL + bcol;
      const double *src = wsrc0 + col0;
      if (full_k && col0 + UTN <= nf) {
#pragma unroll
        for (int j = 0; j < NBMAX / 8; ++j) cp_async16(dst + j * 8 * UCOL, src + (size_t)j * 8 * ld, true);
      } else {
        for (int j = 0; j < kpad / 8; ++j) {
          const bool ok = brow + 8 * j < kw && col0 + bcol < nf;
          cp_async16(dst + j * 8 * UCOL, ok ? src + (size_t)j * 8 * ld : F.A, ok);
        }
      }
      double *cd = sC + ccol * UCP + crow;
      const double *cs = csrc0 + (size_t)col0 * ld;
      if (UPD_PROBE(1)) {
      } else if (row0 >= col0 + UTN - 1 && row0 + UT <= nf) {   // whole tile on or below the diagonal, inside the front
#pragma unroll
        for (int j = 0; j < UTN / 4; ++j) cp_async16(cd + j * 4 * UCP, cs + (size_t)j * 4 * ld, true);
      } else {
        // rows in pairs: a pair is fetched when its second row can hold an entry on or below the diagonal
#pragma unroll 4
        for (int j = 0; j < UTN / 4; ++j) {
          const int c = col0 + ccol + 4 * j;
          const bool ok = row0 + crow + 1 >= c && row0 + crow < nf && c < nf;
          cp_async16(cd + j * 4 * UCP, ok ? cs + (size_t)j * 4 * ld : F.A, ok);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  stage(tj0, 0);  // group 0 = A panel + first tile

  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps of 32 x 32
  const int g = lane >> 2, q = lane & 3;
  const int rbase = row0 + wm * 32 + g;
  const double *pa = sA + q * UROW + wm * 32 + g;
  const double *pc = sC + (wn * 32 + q * 2) * UCP + wm * 32 + g;

  // the A operand is negated once per strip (acc = C, acc += (-L) W^T, store acc: the summation order of the
  // one-tile kernel), each thread flipping the entries it staged itself
  asm volatile("cp.async.wait_group 0;\n" ::);
  for (int idx = tid; idx < kpad * (UT / 2); idx += UPD_THREADS) {
    double2 *p2 = reinterpret_cast<double2 *>(sA + (idx / (UT / 2)) * UROW + (idx % (UT / 2)) * 2);
    double2 v = *p2;
    v.x = -v.x;
    v.y = -v.y;
    *p2 = v;
  }

  // Results of an interior tile are not stored in its own epilogue: they wait in `outv` and leave four at a time
  // between the DMMA batches of the NEXT tile's k-loop (stores need issue slots only, the loop has fifteen spare per
  // DMMA), so the tensor pipe does not idle behind 32 store instructions per thread and tile.
  double outv[4][4][2];
  double *optr = nullptr;
  auto flush_from = [&](int first) {
#pragma unroll
    for (int ce = 0; ce < 8; ++ce)
      if (ce >= first) {
#pragma unroll
        for (int fm = 0; fm < 4; ++fm) optr[(size_t)((ce >> 1) * 8 + (ce & 1)) * ld + fm * 8] = outv[fm][ce >> 1][ce & 1];
      }
    optr = nullptr;
  };

  for (int t = 0; t < ntj; ++t) {
    const int col0 = (tj0 + t) * UTN;
    const int cbase = col0 + wn * 32 + q * 2;
    const bool active = row0 + wm * 32 + 31 >= col0 + wn * 32 && !skipped(tj0 + t);  // not strictly above the diagonal
    // all 32 x 32 entries inside the front, right of the eliminated columns and on or below the diagonal
    const bool interior = active && row0 + wm * 32 >= col0 + wn * 32 + 31 && col0 + wn * 32 >= kcur && row0 + wm * 32 + 31 < nf;
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();  // tile t's W rows and old C values are visible
    double acc[4][4][2];
    double a0[4], b0[4], a1[4], b1[4];
    const double *pb = sB + (t & 1) * NBMAX * UCOL + q * UCOL + wn * 32 + g;
    if (active) {
#pragma unroll
      for (int fm = 0; fm < 4; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn)
#pragma unroll
          for (int e = 0; e < 2; ++e) acc[fm][fn][e] = pc[(fn * 8 + e) * UCP + fm * 8];
#pragma unroll
      for (int f = 0; f < 4; ++f) {   // first fragments: their latency hides behind the barrier and the staging below
        a0[f] = pa[f * 8];
        b0[f] = pb[f * 8];
      }
    }
    __syncthreads();  // sC is free again, and so is the W buffer of tile t - 1
    if (t + 1 < ntj) stage(tj0 + t + 1, (t + 1) & 1);
    if (!active) {
      if (optr) flush_from(0);
      continue;
    }
#pragma unroll
    for (int it = 0; it < NBMAX / 8; ++it) {
      const int k = it * 8;
      if (k < (UPD_PROBE(8) ? 8 : kpad)) {
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          a1[f] = pa[(k + 4) * UROW + f * 8];
          b1[f] = pb[(k + 4) * UCOL + f * 8];
        }
#pragma unroll
        for (int fm = 0; fm < 4; ++fm)
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a0[fm], b0[fn]);
        if (optr) {
#pragma unroll
          for (int fm = 0; fm < 4; ++fm) optr[(size_t)((it >> 1) * 8 + (it & 1)) * ld + fm * 8] = outv[fm][it >> 1][it & 1];
        }
        if (k + 8 < kpad) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            a0[f] = pa[(k + 8) * UROW + f * 8];
            b0[f] = pb[(k + 8) * UCOL + f * 8];
          }
        }
#pragma unroll
        for (int fm = 0; fm < 4; ++fm)
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a1[fm], b1[fn]);
      }
    }
    if (optr) {
      if (kpad < NBMAX) flush_from(kpad / 8);   // a narrow (last) panel ran fewer batches than there are columns
      optr = nullptr;
    }
    if (UPD_PROBE(2)) {
      double sum = 0.0;
#pragma unroll
      for (int fm = 0; fm < 4; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) sum += acc[fm][fn][0] + acc[fm][fn][1];
      if (sum == 1.2345e300) F.A[0] = sum;
    } else if (interior) {
      optr = F.A + rbase + (size_t)cbase * ld;
#pragma unroll
      for (int fm = 0; fm < 4; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) {
          outv[fm][fn][0] = acc[fm][fn][0];
          outv[fm][fn][1] = acc[fm][fn][1];
        }
    } else {
#pragma unroll
      for (int fm = 0; fm < 4; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = rbase + fm * 8, c = cbase + fn * 8 + e;
            if (r < nf && c >= kcur && r >= c) F.A[r + (size_t)c * ld] = acc[fm][fn][e];
          }
    }
  }
  if (optr) flush_from(0);
}

// ---------------------------------------------------------------------------------------------
// Inertia from the pivots: 1x1 by sign, 2x2 by determinant / trace.  out = {pos, neg, zero}.
// ---------------------------------------------------------------------------------------------
__global__ void front_inertia_kernel(const Front *__restrict__ fronts, unsigned long long *out) {
  const Front F = fronts[blockIdx.x];
  int pos = 0, neg = 0, zero = 0;
  for (int k = threadIdx.x; k < F.n; k += blockDim.x) {
    const int b = F.bsz[k];
    if (b == 1) {
      const double d = F.A[k + (size_t)k * F.ld];
      pos += d > 0.0;
      neg += d < 0.0;
      zero += !(d > 0.0) && !(d < 0.0);
    } else if (b == 2) {
      const double a = F.A[k + (size_t)k * F.ld], o = F.A[k + 1 + (size_t)k * F.ld],
                   c = F.A[k + 1 + (size_t)(k + 1) * F.ld];
      // det = a*c - o^2, evaluated scaled by o^2 as in the factorisation
      const double det = (a / o) * (c / o) - 1.0;
      if (det < 0.0) { pos += 1; neg += 1; }
      else if (det > 0.0) { if (a + c > 0.0) pos += 2; else neg += 2; }
      else { zero += 1; if (a + c > 0.0) pos += 1; else if (a + c < 0.0) neg += 1; else zero += 1; }
    }
  }
  __shared__ int s[3];
  if (threadIdx.x < 3) s[threadIdx.x] = 0;
  __syncthreads();
  atomicAdd(&s[0], pos);
  atomicAdd(&s[1], neg);
  atomicAdd(&s[2], zero);
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(&out[threadIdx.x], (unsigned long long)s[threadIdx.x]);
}

}  // namespace ppb
