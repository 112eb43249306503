// Assembly, Schur gather and the triangular solves on fronts.
//
// The solve uses the border rows of L so that one forward and one backward sweep per front do the
// work of the reference's two leaf solves per block (explicit_schur_complement.py:141-153):
//   forward :  v = [P r ; 0],  v <- L^-1 v      =>  v[0:n] = z,  v[n:] = -L_A z = -(A K^-1 r)[rows]
//   diagonal:  w = D^-1 z
//   backward:  v = [w ; x_c[rows]],  v[0:n] <- L^-T-sweep  =>  x = P^T v[0:n] = K^-1 (r - A^T x_c)
#pragma once
#include <cooperative_groups.h>
#include "front.cuh"

namespace ppb {

// ---- assembly -------------------------------------------------------------------------------
// Deterministic scatter-add of the input values into the (zeroed) front arena: destination u gets
// the sum of its sources in input order (duplicates are legal in the KKT, interface.py:454-456).
__global__ void assemble_kernel(const double *__restrict__ vals, const int64_t *__restrict__ dst,
                                const int64_t *__restrict__ ptr, const int64_t *__restrict__ src,
                                int64_t nuniq, double *__restrict__ arena) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nuniq) return;
  double s = 0.0;
  for (int64_t p = ptr[u]; p < ptr[u + 1]; ++p) s += vals[src[p]];
  arena[dst[u]] = s;
}

__global__ void reset_fronts_kernel(const Front *__restrict__ fronts, unsigned long long *inertia) {
  const Front F = fronts[blockIdx.x];
  for (int i = threadIdx.x; i < F.nb; i += blockDim.x) {
    F.perm[i] = i;
    F.ipiv[i] = i;
    F.bsz[i] = 1;
  }
  if (threadIdx.x < 4) F.state[threadIdx.x] = 0;
  if (blockIdx.x == 0 && threadIdx.x < 8 && threadIdx.x != 6 && inertia) inertia[threadIdx.x] = 0ull;  // [7] = finalize ticket
}

// Worst info over a range of fronts -> flag[0] (0 = all fine).
__global__ void collect_info_kernel(const Front *__restrict__ fronts, int count, int *flag) {
  int bad = 0;
  for (int f = threadIdx.x; f < count; f += blockDim.x)
    if (fronts[f].state[ST_INFO] != 0) bad = 1;
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flag[0] = bad;
}

// Tail of the local Schur buffer: [blocks singular?, sparse overflow?, n_pos, n_neg, n_zero, 0, 0, 0] of this rank's
// blocks, so that the ONE all-reduce of the Schur complement also agrees the status across ranks and sums the
// inertia (the reference needs an allgather and three allreduces for that,
// mpi_explicit_schur_complement.py:21,427-429).
__global__ void pack_tail_kernel(double *__restrict__ tail, const int *__restrict__ flag,
                                 const unsigned long long *__restrict__ inertia) {
  const int t = threadIdx.x;
  if (t < 8) {
    double v = 0.0;
    if (t == 0) v = flag[0] ? 1.0 : 0.0;
    else if (t == 1) v = flag[1] ? 1.0 : 0.0;  // sparse path ran out of delayed-pivot capacity (redo densely)
    else if (t >= 2 && t <= 4) v = (double)inertia[t - 2];
    tail[t] = v;
  }
}

// ---- Schur gather ---------------------------------------------------------------------------
// S_local(r,c) = sum over local fronts holding both coupling rows r and c of the front's trailing
// block entry; sources are visited in front order, so the sum is reproducible.  Replaces the
// scatter through sc_data_slices of mpi_explicit_schur_complement.py:249-254,329.
__global__ void schur_gather_kernel(const Front *__restrict__ fronts, const double *__restrict__ arenaA,
                                    const int64_t *__restrict__ src_ptr, const int32_t *__restrict__ src_front,
                                    const int32_t *__restrict__ src_pos, const int64_t *__restrict__ src_aoff,
                                    const int32_t *__restrict__ src_ld, const int64_t *__restrict__ brow_ptr,
                                    const int32_t *__restrict__ brow, int m_c, double *__restrict__ S) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (r >= m_c || r < c) return;
  double s = 0.0;
  const int64_t p0 = src_ptr[r], p1 = src_ptr[r + 1];
  // eight sources in flight: the loads are independent, only the additions keep their order.  src_aoff[p] is
  // the arena offset of (row of r, first border column) in the source front and src_ld[p] its leading dimension
  // -- negated when the front carries only part of the coupling rows and c has to be located first.
  for (int64_t pb = p0; pb < p1; pb += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[q] = 0.0;
      const int64_t p = pb + q;
      if (p >= p1) continue;
      const int ldf = src_ld[p];
      int lo = c;
      if (ldf < 0) {  // partial border: locate c among the front's border rows (c <= r => position <= a)
        const int f = src_front[p], a = src_pos[p];
        const int32_t *br = brow + brow_ptr[f];
        int hi = a;
        lo = 0;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (br[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (br[lo] != c) continue;
      }
      v[q] = arenaA[src_aoff[p] + (int64_t)lo * (ldf < 0 ? -ldf : ldf)];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) s += v[q];
  }
  S[(size_t)r + (size_t)c * m_c] = s;
  S[(size_t)c + (size_t)r * m_c] = s;
}

// The same gather for SMALL coupling systems (few entries, many sources each: 64 scenarios x 50 first-stage variables):
__global__ void schur_gather_warp_kernel(const Front *__restrict__ fronts, const double *__restrict__ arenaA,
                                    const int64_t *__restrict__ src_ptr, const int32_t *__restrict__ src_front,
                                    const int32_t *__restrict__ src_pos, const int64_t *__restrict__ src_aoff,
                                    const int32_t *__restrict__ src_ld, const int64_t *__restrict__ brow_ptr,
                                    const int32_t *__restrict__ brow, int m_c, double *__restrict__ S) {
  // one WARP per entry (r, c) of the lower triangle: lane l adds sources l, l + 32, ... in that order, then a fixed
  // shuffle tree combines the lanes -- every load of an entry is in flight at once (64 scenarios: two per lane) and
  // the summation order is the same in every run.  src_aoff[p] is the arena offset of (row of r, first border column)
  // in the source front and src_ld[p] its leading dimension -- negated when the front carries only part of the
  // coupling rows and c has to be located first.
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int c = blockIdx.y;
  if (r >= m_c || r < c) return;
  double s = 0.0;
  const int64_t p0 = src_ptr[r], p1 = src_ptr[r + 1];
  for (int64_t p = p0 + lane; p < p1; p += 32) {
    const int ldf = src_ld[p];
    int lo = c;
    if (ldf < 0) {  // partial border: locate c among the front's border rows (c <= r => position <= a)
      const int f = src_front[p], a = src_pos[p];
      const int32_t *br = brow + brow_ptr[f];
      int hi = a;
      lo = 0;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (br[mid] < c) lo = mid + 1; else hi = mid;
      }
      if (br[lo] != c) continue;
    }
    s += arenaA[src_aoff[p] + (int64_t)lo * (ldf < 0 ? -ldf : ldf)];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    S[(size_t)r + (size_t)c * m_c] = s;
    S[(size_t)c + (size_t)r * m_c] = s;
  }
}

// The same gather for a SPARSE Schur complement (time-decomposed problems): one thread per pattern entry (r, c), the
// result goes to slot p of the value array -- the role of sc_data_slices in mpi_explicit_schur_complement.py:249-254,
// 326-331, where the all-reduce then moves sc_nnz values instead of m_c^2.
__global__ void schur_gather_sparse_kernel(const double *__restrict__ arenaA, const int64_t *__restrict__ src_ptr,
                                           const int32_t *__restrict__ src_front, const int32_t *__restrict__ src_pos,
                                           const int64_t *__restrict__ src_aoff, const int32_t *__restrict__ src_ld,
                                           const int64_t *__restrict__ brow_ptr, const int32_t *__restrict__ brow,
                                           const int32_t *__restrict__ ent_row, const int32_t *__restrict__ ent_col,
                                           int64_t nnz, double *__restrict__ S) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int r = ent_row[e], c = ent_col[e];
  double s = 0.0;
  for (int64_t p = src_ptr[r]; p < src_ptr[r + 1]; ++p) {  // fronts holding row r, in front order
    const int ldf = src_ld[p];
    int lo = c;
    if (ldf < 0) {
      const int f = src_front[p], a = src_pos[p];
      const int32_t *br = brow + brow_ptr[f];
      int hi = a;
      lo = 0;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (br[mid] < c) lo = mid + 1; else hi = mid;
      }
      if (br[lo] != c) continue;
    }
    s += arenaA[src_aoff[p] + (int64_t)lo * (ldf < 0 ? -ldf : ldf)];
  }
  S[e] = s;
}

// values of the next level: S(p) = reduced Schur sum + the entries of Q that fall on pattern slot p (input order)
__global__ void coupling_values_kernel(const double *__restrict__ Ssum, const double *__restrict__ vals,
                                       const int64_t *__restrict__ q_ptr, const int64_t *__restrict__ q_src,
                                       int64_t nnz, double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  double s = 0.0;
  for (int64_t p = q_ptr[e]; p < q_ptr[e + 1]; ++p) s += vals[q_src[p]];
  out[e] = Ssum[e] + s;
}

// out[j] = a[perm[j]] (+ b[perm[j]]);  out[perm[j]] = in[j]
__global__ void gather_perm_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                   const int32_t *__restrict__ perm, int n, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = b ? a[perm[j]] + b[perm[j]] : a[perm[j]];
}
__global__ void scatter_perm_kernel(const double *__restrict__ in, const int32_t *__restrict__ perm, int n,
                                    double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[perm[j]] = in[j];
}

// coupling front (lower) += reduced Schur sum; Q was assembled into it already.
__global__ void coupling_add_kernel(Front F, const double *__restrict__ Ssum, int m_c) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (r >= m_c || r < c) return;
  F.A[(size_t)r + (size_t)c * F.ld] += Ssum[(size_t)r + (size_t)c * m_c];
}

__global__ void rc_gather_kernel(const double *__restrict__ zarena, const int64_t *__restrict__ src_ptr,
                                 const int64_t *__restrict__ src_boff, int m_c, double *__restrict__ rc) {
  // one warp per coupling row: the sources (one per front that touches the row) are dealt over the lanes, every lane
  // adds its sources in front order, a fixed shuffle tree adds the lanes (reproducible; one round of dependent loads
  // for 64 fronts instead of eight)
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, ln = threadIdx.x & 31;
  if (r >= m_c) return;
  double s = 0.0;
  const int64_t p0 = src_ptr[r], p1 = src_ptr[r + 1];
  for (int64_t pb = p0 + ln; pb < p1; pb += 128) {  // four sources in flight per lane
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = pb + 32 * q < p1 ? zarena[src_boff[pb + 32 * q]] : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += v[q];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (ln == 0) rc[r] = s;
}

__global__ void vec_add_kernel(const double *__restrict__ a, const double *__restrict__ b, int n,
                               double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// ---- residual r = b - K x for iterative refinement ---------------------------------------------
// K is applied from the ORIGINAL input values through row-wise lists (col, source) built at symbolic time;
// every sum runs in list order, block reductions are trees: the result is reproducible.
// Columns < ldim index the local solution, the others the coupling solution.
__global__ void residual_rows_kernel(const long long *__restrict__ ptr, const int *__restrict__ col,
                                     const long long *__restrict__ src, const double *__restrict__ vals,
                                     const double *__restrict__ x, const double *__restrict__ xc, long long ldim,
                                     const double *__restrict__ b, double *__restrict__ r,
                                     double *__restrict__ part, long long nrows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double ri = 0.0, bi = 0.0;
  if (i < nrows) {
    double s = 0.0;
    for (long long p = ptr[i]; p < ptr[i + 1]; ++p) {
      const int c = col[p];
      s += vals[src[p]] * (c < ldim ? x[c] : xc[c - ldim]);
    }
    bi = b[i];
    ri = bi - s;
    r[i] = ri;
  }
  __shared__ double s1[256], s2[256];
  s1[threadIdx.x] = ri * ri;
  s2[threadIdx.x] = bi * bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = s1[0]; part[2 * blockIdx.x + 1] = s2[0]; }
}

// buf[g] = -(sum over local borders A x)(g) for g < m_c;  buf[m_c], buf[m_c+1] = sums of the row partials
__global__ void residual_border_kernel(const long long *__restrict__ ptr, const int *__restrict__ col,
                                       const long long *__restrict__ src, const double *__restrict__ vals,
                                       const double *__restrict__ x, int m_c, const double *__restrict__ part,
                                       int nparts, double *__restrict__ buf) {
  // one warp per coupling row: the row's entries (one per block that touches it) are dealt over the lanes, every lane
  // adds its entries in list order and a fixed shuffle tree adds the lanes -- reproducible, and a row with 64 entries
  // costs one round of dependent loads instead of eight
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, ln = threadIdx.x & 31;
  if (g < m_c) {
    double s = 0.0;
    const long long p0 = ptr[g], p1 = ptr[g + 1];
    for (long long pb = p0 + ln; pb < p1; pb += 128) {  // four products in flight per lane
      double t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q] = pb + 32 * q < p1 ? vals[src[pb + 32 * q]] * x[col[pb + 32 * q]] : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) s += t[q];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (ln == 0) buf[g] = -s;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {
    // the two norms: warp w sums the partials of component w, each lane a strided subset, then a fixed tree
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int k0 = lane; k0 < nparts; k0 += 256) {  // eight loads in flight per lane
      double t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = k0 + 32 * q < nparts ? part[2 * (k0 + 32 * q) + w] : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += t[q];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) buf[m_c + w] = s;
  }
}

// r_c = b_c - Q x_c + bufsum[0:m_c]   (one block);  out2 = {bufsum[m_c] + |r_c|^2, bufsum[m_c+1] + |b_c|^2}
__global__ void residual_coupling_kernel(const long long *__restrict__ ptr, const int *__restrict__ col,
                                         const long long *__restrict__ src, const double *__restrict__ vals,
                                         const double *__restrict__ xc, const double *__restrict__ bc,
                                         const double *__restrict__ bufsum, int m_c, double *__restrict__ rc,
                                         double *__restrict__ out2) {
  __shared__ double s1[256], s2[256];
  double a1 = 0.0, a2 = 0.0;
  for (int g = threadIdx.x; g < m_c; g += blockDim.x) {
    double s = 0.0;
    for (long long p = ptr[g]; p < ptr[g + 1]; ++p) s += vals[src[p]] * xc[col[p]];
    const double r = bc[g] - s + bufsum[g];
    rc[g] = r;
    a1 += r * r;
    a2 += bc[g] * bc[g];
  }
  s1[threadIdx.x] = a1;
  s2[threadIdx.x] = a2;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = bufsum[m_c] + s1[0]; out2[1] = bufsum[m_c + 1] + s2[0]; }
}

__global__ void axpy1_kernel(double *__restrict__ x, const double *__restrict__ d, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += d[i];
}

// ---- triangular solves ----------------------------------------------------------------------
constexpr int SB = 32;         // columns per sweep block
constexpr int SPITCH = SB + 1;

__device__ __forceinline__ void load_tri(const Front &F, int kb, int bw, double *tri) {
  for (int idx = threadIdx.x; idx < bw * bw; idx += blockDim.x) {
    const int c = idx / bw, r = idx % bw;
    double v = 0.0;
    if (r > c && !(r == c + 1 && F.bsz[kb + c] == 2)) v = F.A[(size_t)(kb + r) + (size_t)(kb + c) * F.ld];
    tri[c * SPITCH + r] = v;
  }
}

// v lives in dynamic shared memory (nf doubles) followed by the SB x SPITCH diagonal block.
// forward sweep of one dense front; `r` (and `r2`, optional second addend) indexed by original row
template <int NT>
__device__ __forceinline__ void front_forward_body(const Front &F, double *sm, const double *__restrict__ r,
                                                   const double *__restrict__ r2) {
  const int n = F.n, nf = F.nf, ld = F.ld, tid = threadIdx.x;
  double *v = sm;
  double *tri = sm + ((nf + 1) & ~1);
  for (int i = tid; i < nf; i += NT) {
    double val = 0.0;
    if (i < n) {
      const int o = F.perm[i];
      val = r2 ? r[o] + r2[o] : r[o];
    }
    v[i] = val;
  }
  __syncthreads();
  for (int kb = 0; kb < n; kb += SB) {
    const int bw = min(SB, n - kb);
    load_tri(F, kb, bw, tri);
    __syncthreads();
    if (tid < 32) {
      double x = tid < bw ? v[kb + tid] : 0.0;
      for (int c = 0; c < bw; ++c) {
        const double xc = __shfl_sync(0xffffffffu, x, c);
        if (tid > c && tid < bw) x -= tri[c * SPITCH + tid] * xc;
      }
      if (tid < bw) v[kb + tid] = x;
    }
    __syncthreads();
    const bool straddle = F.bsz[kb + bw - 1] == 2;  // 2x2 pivot across the block edge
    for (int i = kb + bw + tid; i < nf; i += NT) {
      const double *__restrict__ Li = F.A + i + (size_t)kb * ld;
      double acc = v[i];
      const int cend = (straddle && i == kb + bw) ? bw - 1 : bw;
#pragma unroll 4
      for (int c = 0; c < cend; ++c) acc -= Li[(size_t)c * ld] * v[kb + c];
      v[i] = acc;
    }
    __syncthreads();
  }
  // w = D^-1 z
  for (int k = tid; k < n; k += NT) {
    const int b = F.bsz[k];
    if (b == 1) {
      const double d = F.A[k + (size_t)k * ld];
      F.zbuf[k] = d != 0.0 ? v[k] / d : 0.0;
    } else if (b == 2) {
      const double e21 = F.A[k + 1 + (size_t)k * ld];
      const double akm1 = F.A[k + (size_t)k * ld] / e21, ak = F.A[k + 1 + (size_t)(k + 1) * ld] / e21;
      const double denom = akm1 * ak - 1.0;
      const double bkm1 = v[k] / e21, bk = v[k + 1] / e21;
      F.zbuf[k] = (ak * bkm1 - bk) / denom;
      F.zbuf[k + 1] = (akm1 * bk - bkm1) / denom;
    }
  }
  for (int a = tid; a < F.m; a += NT) F.bvec[a] = v[F.nb + a];
}

template <int NT>
__global__ void __launch_bounds__(NT) front_forward_kernel(const Front *__restrict__ fronts,
                                                           const double *__restrict__ rhs,
                                                           const int64_t *__restrict__ rhs_off) {
  extern __shared__ double sm[];
  const Front F = fronts[blockIdx.x];
  front_forward_body<NT>(F, sm, rhs + rhs_off[blockIdx.x], nullptr);
}

// backward sweep of one dense front; border values from xc[br[a]], result scattered to out[perm[i]]
template <int NT>
__device__ __forceinline__ void front_backward_body(const Front &F, double *sm, const double *__restrict__ xc,
                                                    const int32_t *__restrict__ br, double *__restrict__ out) {
  const int n = F.n, nf = F.nf, ld = F.ld, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  double *v = sm;
  double *tri = sm + ((nf + 1) & ~1);
  for (int i = tid; i < F.nb; i += NT) v[i] = i < n ? F.zbuf[i] : 0.0;
  if (F.m > 0)
    for (int a = tid; a < F.m; a += NT) v[F.nb + a] = xc[br[a]];
  __syncthreads();
  for (int kb = ((n - 1) / SB) * SB; kb >= 0; kb -= SB) {
    const int bw = min(SB, n - kb);
    load_tri(F, kb, bw, tri);
    const bool straddle = F.bsz[kb + bw - 1] == 2;
    // v[kb+c] -= L(kb+bw:nf, kb+c)^T v[kb+bw:nf]   (one warp per column, coalesced down the column)
    for (int c = warp; c < bw; c += NT / 32) {
      const double *__restrict__ Lc = F.A + (size_t)(kb + c) * ld;
      const int ibeg = kb + bw + ((straddle && c == bw - 1) ? 1 : 0);
      double acc = 0.0;
      for (int i = ibeg + lane; i < nf; i += 32) acc += Lc[i] * v[i];
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) v[kb + c] -= acc;
    }
    __syncthreads();
    if (tid < 32) {
      double y = tid < bw ? v[kb + tid] : 0.0;
      for (int rr = bw - 1; rr > 0; --rr) {
        const double xr = __shfl_sync(0xffffffffu, y, rr);
        if (tid < rr) y -= tri[tid * SPITCH + rr] * xr;
      }
      if (tid < bw) v[kb + tid] = y;
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += NT) out[F.perm[i]] = v[i];
}

template <int NT>
__global__ void __launch_bounds__(NT) front_backward_kernel(const Front *__restrict__ fronts,
                                                            const double *__restrict__ xc,
                                                            const int64_t *__restrict__ brow_ptr,
                                                            const int32_t *__restrict__ brow,
                                                            double *__restrict__ x,
                                                            const int64_t *__restrict__ x_off) {
  extern __shared__ double sm[];
  const Front F = fronts[blockIdx.x];
  front_backward_body<NT>(F, sm, xc, brow + brow_ptr[blockIdx.x], x + x_off[blockIdx.x]);
}

// The coupling system S x_c = r_c + sum_i rc_i in ONE launch (one CTA): right-hand side formed on the fly, forward
// sweep, D^-1, backward sweep.  The coupling front has no border rows.
template <int NT>
__global__ void __launch_bounds__(NT) coupling_solve_kernel(const Front *__restrict__ front,
                                                            const double *__restrict__ rhs_c,
                                                            const double *__restrict__ rc_sum,
                                                            double *__restrict__ x_c) {
  extern __shared__ double sm[];
  const Front F = *front;
  front_forward_body<NT>(F, sm, rhs_c, rc_sum);
  __syncthreads();  // zbuf written above is read below by other threads of this CTA
  front_backward_body<NT>(F, sm, nullptr, nullptr, x_c);
}

// ---- the same sweeps for TALL fronts, one thread-block cluster per front ------------------------------------
// One CTA streaming the 67 MB factor of a 4 082-row root alone gets ~18 GB/s (3.7 ms per sweep, 2 % of the HBM rate,
// VERDICT r1); when there are fewer tall fronts than SMs the rows are dealt over a cluster instead: 32-row tiles,
// tile t belongs to CTA t mod C, every CTA keeps the running vector of its own rows in shared memory (about one row
// per thread at 8 CTAs x 512 threads and 4 082 rows).
//   forward, step s (columns 32 s ...): the owner of tile s solves the 32 x 32 diagonal block in one warp and writes
//     the 32 solved values into every CTA's shared memory (DSMEM); after ONE cluster barrier every CTA updates its rows.
//   backward, step s: every CTA forms the partial dot products of its rows with the 32 columns and writes them into
//     the owner's shared memory; after ONE cluster barrier the owner adds them in rank order (reproducible), solves
//     the diagonal block and keeps the result -- the rows of tile s are its own.
constexpr int CS_NT = 512, CS_NWARP = CS_NT / 32, CS_MAXC = 8;
#ifdef PP_TRACE_SOLVE
__device__ long long g_cs_trace[8];
#define CS_TR(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_cs_trace[k] += clock64() - t_last, t_last = clock64(); } while (0)
#else
#define CS_TR(k) do { } while (0)
#endif

struct ClusterSolve {
  double *vloc;   // own tiles, 32 doubles each: tile t = lt * C + rank at vloc[32 lt]
  double *xs;     // [2][32] solved block of the current forward step (double-buffered by step parity)
  double *part;   // [2][CS_MAXC][32] partial dot products of a backward step, one slot per source CTA
  double *tri;    // [32][SPITCH] diagonal block
  int ntile_loc;
};

__host__ __device__ inline size_t cluster_solve_smem(int nf, int C) {
  const int ntile = (nf + 31) / 32, loc = (ntile + C - 1) / C + 1;
  return ((size_t)loc * 32 + 2 * 32 + 2 * CS_MAXC * 32 + SB * SPITCH) * sizeof(double);
}

__device__ __forceinline__ ClusterSolve carve_cluster_solve(double *sm, int nf, int C) {
  ClusterSolve S;
  const int ntile = (nf + 31) / 32;
  S.ntile_loc = (ntile + C - 1) / C + 1;
  S.vloc = sm;
  S.xs = S.vloc + (size_t)S.ntile_loc * 32;
  S.part = S.xs + 64;
  S.tri = S.part + 2 * CS_MAXC * 32;
  return S;
}

// v <- L^-1 v on the rows of this CTA; v holds [P r ; 0] on entry
__device__ __forceinline__ void cluster_forward_sweep(cooperative_groups::cluster_group &cl, const Front &F,
                                                      const ClusterSolve &S) {
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const int n = F.n, nf = F.nf, ld = F.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntile = (nf + 31) / 32;
  const int nstep = (n + 31) / 32;
#ifdef PP_TRACE_SOLVE
  long long t_last = clock64();
#endif
  for (int s = 0; s < nstep; ++s) {
    const int kb = 32 * s, bw = min(32, n - kb), p = s & 1;
    const int last_kind = F.bsz[kb + bw - 1];   // fetched ahead of the barrier that precedes its use
    CS_TR(0);
    if (s % C == rank) {
      load_tri(F, kb, bw, S.tri);
      __syncthreads();
      if (warp == 0) {
        double x = lane < bw ? S.vloc[32 * (s / C) + lane] : 0.0;
        for (int c = 0; c < bw; ++c) {
          const double xc = __shfl_sync(0xffffffffu, x, c);
          if (lane > c && lane < bw) x -= S.tri[c * SPITCH + lane] * xc;
        }
        if (lane < bw) S.vloc[32 * (s / C) + lane] = x;
        for (int q = 0; q < C; ++q) {   // the solved block goes to every CTA of the cluster
          double *dst = cl.map_shared_rank(S.xs, q);
          dst[32 * p + lane] = lane < bw ? x : 0.0;
        }
      }
    }
    CS_TR(1);
    cl.sync();
    CS_TR(2);
    // rows below the block, own tiles only: tile t = lt * C + rank > s
    const bool straddle = last_kind == 2;  // 2x2 pivot across the block edge
    const double *xs = S.xs + 32 * p;
    int lt0 = (s - rank + C - 1) / C;   // first own tile >= s (the rows of tile s beyond a partial block are below it)
    if (lt0 < 0) lt0 = 0;
    for (int lt = lt0 + warp; lt * C + rank < ntile; lt += CS_NWARP) {
      const int i = 32 * (lt * C + rank) + lane;
      if (i >= nf || i < kb + bw) continue;
      const double *__restrict__ Li = F.A + i + (size_t)kb * ld;
      double acc = S.vloc[32 * lt + lane];
      const int cend = (straddle && i == kb + bw) ? bw - 1 : bw;
      double l[32];   // every load of the row's 32 entries is issued before the first use
#pragma unroll
      for (int c = 0; c < 32; ++c) l[c] = c < cend ? Li[(size_t)c * ld] : 0.0;
#pragma unroll
      for (int c = 0; c < 32; ++c) acc -= l[c] * xs[c];
      S.vloc[32 * lt + lane] = acc;
    }
    // no barrier here: the next step's owner only touches its own tile (updated by its own warps: block barrier
    // below), and xs is double-buffered
    CS_TR(3);
    __syncthreads();
    CS_TR(4);
  }
}

// v <- L^-T-sweep on the rows of this CTA; on entry v = [D^-1 z ; border values]
__device__ __forceinline__ void cluster_backward_sweep(cooperative_groups::cluster_group &cl, const Front &F,
                                                       const ClusterSolve &S) {
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const int n = F.n, nf = F.nf, ld = F.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntile = (nf + 31) / 32;
  const int nstep = (n + 31) / 32;
  __shared__ double wpart[CS_NWARP][33];
  for (int s = nstep - 1; s >= 0; --s) {
    const int kb = 32 * s, bw = min(32, n - kb), p = s & 1, owner = s % C;
    const bool straddle = F.bsz[kb + bw - 1] == 2;
    // partial dot products over own rows below the block: warp w walks its tiles, lane = row in tile; for every
    // column c the 32 rows of a tile are reduced with shuffles, the warp keeps one running sum per column (lane c)
    double acc[32];   // this lane's row against the 32 columns, summed over the warp's tiles
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0;
    int lt0 = (s - rank + C - 1) / C;   // first own tile >= s
    if (lt0 < 0) lt0 = 0;
    for (int lt = lt0 + warp; lt * C + rank < ntile; lt += CS_NWARP) {
      const int i = 32 * (lt * C + rank) + lane;
      if (i >= nf || i < kb + bw) continue;
      const double vi = S.vloc[32 * lt + lane];
      const double *__restrict__ Li = F.A + i + (size_t)kb * ld;
      const int cend = (straddle && i == kb + bw) ? bw - 1 : bw;
      double l[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) l[c] = c < cend ? Li[(size_t)c * ld] : 0.0;
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += l[c] * vi;
    }
    // transpose-reduce: after the five rounds lane c holds the sum over the lanes of acc[c] (31 shuffles)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int k = 0; k < o; ++k) {
        const double send = upper ? acc[k] : acc[k + o];
        const double keep = upper ? acc[k + o] : acc[k];
        acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    const double colsum = acc[0];
    wpart[warp][lane] = colsum;
    __syncthreads();
    if (warp == 0) {   // warps in index order, then to the owner's slot of this CTA
      double t = 0.0;
      for (int w = 0; w < CS_NWARP; ++w) t += wpart[w][lane];
      double *dst = cl.map_shared_rank(S.part, owner);
      dst[(p * CS_MAXC + rank) * 32 + lane] = t;
    }
    cl.sync();
    if (rank == owner) {
      load_tri(F, kb, bw, S.tri);
      __syncthreads();
      if (warp == 0) {
        double y = lane < bw ? S.vloc[32 * (s / C) + lane] : 0.0;
        for (int q = 0; q < C; ++q) y -= S.part[(p * CS_MAXC + q) * 32 + lane];   // rank order
        for (int rr = bw - 1; rr > 0; --rr) {
          const double xr = __shfl_sync(0xffffffffu, y, rr);
          if (lane < rr) y -= S.tri[lane * SPITCH + rr] * xr;
        }
        if (lane < bw) S.vloc[32 * (s / C) + lane] = y;
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void cluster_load_rhs(const Front &F, const ClusterSolve &S, int C, int rank,
                                                 const double *__restrict__ r, const double *__restrict__ r2) {
  const int ntile = (F.nf + 31) / 32;
  for (int idx = threadIdx.x; idx < S.ntile_loc * 32; idx += CS_NT) {
    const int lt = idx >> 5, i = 32 * (lt * C + rank) + (idx & 31);
    double val = 0.0;
    if (lt * C + rank < ntile && i < F.n) {
      const int o = F.perm[i];
      val = r2 ? r[o] + r2[o] : r[o];
    }
    S.vloc[idx] = val;
  }
}

// z -> global scratch (zbuf / bvec), then w = D^-1 z in place (pairs of a 2x2 pivot are handled by one thread)
__device__ __forceinline__ void cluster_store_forward(cooperative_groups::cluster_group &cl, const Front &F,
                                                      const ClusterSolve &S, int C, int rank) {
  const int ntile = (F.nf + 31) / 32, n = F.n, ld = F.ld;
  for (int idx = threadIdx.x; idx < S.ntile_loc * 32; idx += CS_NT) {
    const int lt = idx >> 5, i = 32 * (lt * C + rank) + (idx & 31);
    if (lt * C + rank >= ntile || i >= F.nf) continue;
    if (i < n) F.zbuf[i] = S.vloc[idx];
    else if (i >= F.nb) F.bvec[i - F.nb] = S.vloc[idx];
  }
  __threadfence();
  cl.sync();
  for (int k = rank * CS_NT + threadIdx.x; k < n; k += C * CS_NT) {
    const int b = F.bsz[k];
    if (b == 1) {
      const double d = F.A[k + (size_t)k * ld];
      F.zbuf[k] = d != 0.0 ? F.zbuf[k] / d : 0.0;
    } else if (b == 2) {
      const double e21 = F.A[k + 1 + (size_t)k * ld];
      const double akm1 = F.A[k + (size_t)k * ld] / e21, ak = F.A[k + 1 + (size_t)(k + 1) * ld] / e21;
      const double denom = akm1 * ak - 1.0;
      const double bkm1 = F.zbuf[k] / e21, bk = F.zbuf[k + 1] / e21;
      F.zbuf[k] = (ak * bkm1 - bk) / denom;
      F.zbuf[k + 1] = (akm1 * bk - bkm1) / denom;
    }
  }
}

__device__ __forceinline__ void cluster_load_backward(const Front &F, const ClusterSolve &S, int C, int rank,
                                                      const double *__restrict__ xc, const int32_t *__restrict__ br) {
  const int ntile = (F.nf + 31) / 32;
  for (int idx = threadIdx.x; idx < S.ntile_loc * 32; idx += CS_NT) {
    const int lt = idx >> 5, i = 32 * (lt * C + rank) + (idx & 31);
    double val = 0.0;
    if (lt * C + rank < ntile && i < F.nf) {
      if (i < F.n) val = F.zbuf[i];
      else if (i >= F.nb && F.m > 0) val = xc[br[i - F.nb]];
    }
    S.vloc[idx] = val;
  }
}

__device__ __forceinline__ void cluster_store_backward(const Front &F, const ClusterSolve &S, int C, int rank,
                                                       double *__restrict__ out) {
  const int ntile = (F.nf + 31) / 32;
  for (int idx = threadIdx.x; idx < S.ntile_loc * 32; idx += CS_NT) {
    const int lt = idx >> 5, i = 32 * (lt * C + rank) + (idx & 31);
    if (lt * C + rank < ntile && i < F.n) out[F.perm[i]] = S.vloc[idx];
  }
}

__global__ void __launch_bounds__(CS_NT) front_forward_cluster_kernel(const Front *__restrict__ fronts,
                                                                      const double *__restrict__ rhs,
                                                                      const int64_t *__restrict__ rhs_off) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank(), f = blockIdx.x / C;
  extern __shared__ double sm[];
  const Front F = fronts[f];
  const ClusterSolve S = carve_cluster_solve(sm, F.nf, C);
  cluster_load_rhs(F, S, C, rank, rhs + rhs_off[f], nullptr);
  __syncthreads();
  cl.sync();   // every CTA's shared memory is set up before anybody writes into it remotely
  cluster_forward_sweep(cl, F, S);
  cluster_store_forward(cl, F, S, C, rank);
}

__global__ void __launch_bounds__(CS_NT) front_backward_cluster_kernel(const Front *__restrict__ fronts,
                                                                       const double *__restrict__ xc,
                                                                       const int64_t *__restrict__ brow_ptr,
                                                                       const int32_t *__restrict__ brow,
                                                                       double *__restrict__ x,
                                                                       const int64_t *__restrict__ x_off) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank(), f = blockIdx.x / C;
  extern __shared__ double sm[];
  const Front F = fronts[f];
  const ClusterSolve S = carve_cluster_solve(sm, F.nf, C);
  cluster_load_backward(F, S, C, rank, xc, brow + brow_ptr[f]);
  __syncthreads();
  cl.sync();
  cluster_backward_sweep(cl, F, S);
  cluster_store_backward(F, S, C, rank, x + x_off[f]);
  cl.sync();   // nobody leaves while its shared memory may still be written remotely
}

// S x_c = r_c + rc_sum for a tall coupling front: forward, D^-1, backward in one cluster
__global__ void __launch_bounds__(CS_NT) coupling_solve_cluster_kernel(const Front *__restrict__ front,
                                                                       const double *__restrict__ rhs_c,
                                                                       const double *__restrict__ rc_sum,
                                                                       double *__restrict__ x_c) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  extern __shared__ double sm[];
  const Front F = *front;
  const ClusterSolve S = carve_cluster_solve(sm, F.nf, C);
  cluster_load_rhs(F, S, C, rank, rhs_c, rc_sum);
  __syncthreads();
  cl.sync();
  cluster_forward_sweep(cl, F, S);
  cluster_store_forward(cl, F, S, C, rank);
  __threadfence();
  cl.sync();   // D^-1 z of every CTA is in zbuf
  cluster_load_backward(F, S, C, rank, nullptr, nullptr);
  __syncthreads();
  cl.sync();
  cluster_backward_sweep(cl, F, S);
  cluster_store_backward(F, S, C, rank, x_c);
  cl.sync();
}

}  // namespace ppb
