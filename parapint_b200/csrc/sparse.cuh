// Multifrontal factorisation / solves of the SUBTREE part of every diagonal block (one CTA per block).
//
// The CTA walks its block's assembly tree in postorder.  Each small front is assembled in shared
// memory (original entries + the contribution blocks of its children, popped from a per-block stack
// in HBM), partially factorised with threshold pivoting among its fully-summed rows (1x1 and 2x2
// pivots; columns that find no acceptable pivot are DELAYED to the parent, MA57-style), its L/D
// columns are written to the block's factor arena and its contribution block is pushed for the
// parent -- or, for children of the root, added into the block's dense root front, which the batched
// Bunch-Kaufman kernels of factor.cuh then finish (they also see the delayed columns, so no pivot is
// ever forced: inertia stays exact).
#pragma once
#include "front.cuh"

namespace ppb {

struct PlanDev {
  int n, m, nT, DR, ns, pad;
  const int *rootcols;
  const int *col_ptr, *cols, *row_ptr, *rows, *rel, *parent, *nchild, *dcap, *fid_off, *fs_off;
  const long long *l_off;
  const int *ent_ptr, *tgt_row, *tgt_col, *tgt_src_ptr, *tgt_src;
};

struct SparseBlock {
  int plan, root;        // plan index; index of the dense root front
  long long val_off;     // base index of this block's input values
  double *L;             // factor arena of the subtree supernodes
  double *stack;         // contribution-block stack
  long long stack_cap;
  int *fid;              // per supernode: original ids of the front rows after pivoting
  int *pbz;              // per supernode: pivot flags (1, 2, 0) of the eliminated columns
  int *meta;             // per supernode: {ne, S}
  int *rootids;          // [nT + DR] original id per root position, -1 for unused delayed slots
  int *info;             // [0] overflow / failure flag, [1] delayed pivots that reached the root
};

constexpr int SF_NT = 128;           // threads per CTA in the subtree kernels
constexpr int SF_SBUF = 96;          // largest front held in shared memory
constexpr int SF_LDF = SF_SBUF + 1;  // odd pitch: conflict-free row and column walks
constexpr size_t SF_SMEM = (size_t)SF_SBUF * SF_LDF * sizeof(double) + 3 * SF_SBUF * sizeof(double) +
                           4 * SF_SBUF * sizeof(int) + 64;

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void warp_argmax(double &v, int &i) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
    if (i2 >= 0 && (i < 0 || v2 > v || (v2 == v && i2 < i))) { v = v2; i = i2; }
  }
}

__device__ __forceinline__ double &fent(double *F, int r, int c) {  // lower-triangle accessor
  return r >= c ? F[r + c * SF_LDF] : F[c + r * SF_LDF];
}

// symmetric interchange of positions a < b in the lower-stored front (L rows included)
__device__ __forceinline__ void front_swap(double *F, int S, int a, int b, int *fid) {
  if (a == b) return;
  for (int i = threadIdx.x; i < S; i += SF_NT) {
    double *p, *q;
    if (i < a) { p = &F[a + i * SF_LDF]; q = &F[b + i * SF_LDF]; }
    else if (i == a) { p = &F[a + a * SF_LDF]; q = &F[b + b * SF_LDF]; }
    else if (i < b) { p = &F[i + a * SF_LDF]; q = &F[b + i * SF_LDF]; }
    else if (i == b) continue;
    else { p = &F[i + a * SF_LDF]; q = &F[i + b * SF_LDF]; }
    const double t = *p;
    *p = *q;
    *q = t;
  }
  if (threadIdx.x == 0) { const int t = fid[a]; fid[a] = fid[b]; fid[b] = t; }
  __syncthreads();
}

// Partial factorisation of the S x S front in shared memory; the first fs rows are fully summed.
// Returns the number of eliminated columns; inertia counts are accumulated by thread 0 in cnt[3].
__device__ int factor_front(double *F, int S, int fs, int *fid, int *bsz, double *w0, double *w1, double u,
                            double pivtol, int *cnt, int *sh_i) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = SF_NT / 32;
  int t = 0;
  while (t < fs) {
    if (warp == 0) {
      int kind = 0, pc = -1, pr = -1;
      for (int c = t; c < fs; ++c) {
        double cmax = 0.0, fbest = -1.0;
        int r = -1;
        for (int i = t + lane; i < S; i += 32) {
          if (i == c) continue;
          const double v = fabs(fent(F, i, c));
          cmax = fmax(cmax, v);
          if (i < fs && v > fbest) { fbest = v; r = i; }
        }
        cmax = warp_max(cmax);
        warp_argmax(fbest, r);
        const double dcc = F[c + c * SF_LDF];
        if (fabs(dcc) > pivtol && fabs(dcc) >= u * cmax) { kind = 1; pc = c; break; }
        if (r >= 0 && fbest > pivtol) {
          double cm_c = 0.0, cm_r = 0.0;
          for (int i = t + lane; i < S; i += 32) {
            if (i == c || i == r) continue;
            cm_c = fmax(cm_c, fabs(fent(F, i, c)));
            cm_r = fmax(cm_r, fabs(fent(F, i, r)));
          }
          cm_c = warp_max(cm_c);
          cm_r = warp_max(cm_r);
          const double drr = F[r + r * SF_LDF], b = fent(F, r, c);
          const double det = dcc * drr - b * b;
          const double g1 = (fabs(drr) * cm_c + fabs(b) * cm_r) / fabs(det);
          const double g2 = (fabs(b) * cm_c + fabs(dcc) * cm_r) / fabs(det);
          if (g1 * u <= 1.0 && g2 * u <= 1.0) { kind = 2; pc = c; pr = r; break; }
          if (fabs(drr) > pivtol && fabs(drr) >= u * fmax(cm_r, fabs(b))) { kind = 1; pc = r; break; }
        }
      }
      if (lane == 0) { sh_i[0] = kind; sh_i[1] = pc; sh_i[2] = pr; }
    }
    __syncthreads();
    const int kind = sh_i[0];
    int pc = sh_i[1], pr = sh_i[2];
    __syncthreads();
    if (kind == 0) break;
    if (kind == 2 && pr < pc) { const int x = pc; pc = pr; pr = x; }
    front_swap(F, S, t, pc, fid);
    if (kind == 1) {
      const double d = F[t + t * SF_LDF];
      for (int i = t + 1 + tid; i < S; i += SF_NT) w0[i] = F[i + t * SF_LDF];
      __syncthreads();
      const double rd = 1.0 / d;
      for (int j = t + 1 + warp; j < S; j += NW) {
        const double wj = w0[j] * rd;
        for (int i = j + lane; i < S; i += 32) F[i + j * SF_LDF] -= w0[i] * wj;
      }
      for (int i = t + 1 + tid; i < S; i += SF_NT) F[i + t * SF_LDF] = w0[i] * rd;
      if (tid == 0) {
        bsz[t] = 1;
        cnt[d > 0.0 ? 0 : (d < 0.0 ? 1 : 2)]++;
      }
      __syncthreads();
      t += 1;
    } else {
      front_swap(F, S, t + 1, pr, fid);
      const double e11 = F[t + t * SF_LDF], e21 = F[t + 1 + t * SF_LDF], e22 = F[t + 1 + (t + 1) * SF_LDF];
      for (int i = t + 2 + tid; i < S; i += SF_NT) {
        w0[i] = F[i + t * SF_LDF];
        w1[i] = F[i + (t + 1) * SF_LDF];
      }
      __syncthreads();
      const double d11 = e22 / e21, d22 = e11 / e21;
      const double sc = (1.0 / (d11 * d22 - 1.0)) / e21;
      for (int j = t + 2 + warp; j < S; j += NW) {
        const double a0 = w0[j], a1 = w1[j];
        for (int i = j + lane; i < S; i += 32) {
          const double l0 = sc * (d11 * w0[i] - w1[i]), l1 = sc * (d22 * w1[i] - w0[i]);
          F[i + j * SF_LDF] -= l0 * a0 + l1 * a1;
        }
      }
      for (int i = t + 2 + tid; i < S; i += SF_NT) {
        F[i + t * SF_LDF] = sc * (d11 * w0[i] - w1[i]);
        F[i + (t + 1) * SF_LDF] = sc * (d22 * w1[i] - w0[i]);
      }
      if (tid == 0) {
        bsz[t] = 2;
        bsz[t + 1] = 0;
        const double det = d11 * d22 - 1.0;  // sign of the determinant (scaled by e21^2 > 0)
        if (det < 0.0) { cnt[0]++; cnt[1]++; }
        else if (det > 0.0) { if (e11 + e22 > 0.0) cnt[0] += 2; else cnt[1] += 2; }
        else { cnt[2]++; cnt[e11 + e22 > 0.0 ? 0 : 1]++; }
      }
      __syncthreads();
      t += 2;
    }
  }
  return t;
}

// record layout on the contribution stack (doubles):  [ids: dim ints, padded][dim*dim matrix][footer 2]
// footer = {dim | nd_out<<16 | supernode<<32 ... } stored as three ints + padding in two doubles
__device__ __forceinline__ long long rec_size(int dim) {
  return (long long)((dim + 1) / 2) + (long long)dim * dim + 2;
}

__global__ void __launch_bounds__(SF_NT) subtree_factor_kernel(const SparseBlock *__restrict__ blocks,
                                                               const PlanDev *__restrict__ plans,
                                                               const Front *__restrict__ fronts,
                                                               const double *__restrict__ vals, double u,
                                                               double pivtol, unsigned long long *inertia) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  double *F = reinterpret_cast<double *>(sm_raw);
  double *w0 = F + SF_SBUF * SF_LDF;
  double *w1 = w0 + SF_SBUF;
  int *fid = reinterpret_cast<int *>(w1 + 2 * SF_SBUF);  // (one spare row of doubles keeps alignment simple)
  int *bsz = fid + SF_SBUF;
  int *map = bsz + SF_SBUF;
  int *sh_i = map + SF_SBUF;  // 16 ints of scratch

  const SparseBlock B = blocks[blockIdx.x];
  const PlanDev P = plans[B.plan];
  const Front R = fronts[B.root];
  const int tid = threadIdx.x;
  int *cnt = sh_i + 8;
  if (tid < 3) cnt[tid] = 0;
  long long sp = 0;  // stack pointer (doubles)
  int ndroot = 0;
  bool failed = false;

  for (int s = 0; s < P.ns && !failed; ++s) {
    const int c0 = P.col_ptr[s], nc = P.col_ptr[s + 1] - c0;
    const int r0 = P.row_ptr[s], ncb = P.row_ptr[s + 1] - r0;
    const int nch = P.nchild[s];
    // ---- incoming delayed pivots: walk the children's footers from the top of the stack ----
    int nd_in = 0;
    {
      long long q = sp;
      for (int c = 0; c < nch; ++c) {
        const int *ft = reinterpret_cast<const int *>(B.stack + q - 2);
        nd_in += ft[1];
        q -= rec_size(ft[0]);
      }
    }
    const int fs = nc + nd_in, S = fs + ncb;
    if (S > SF_SBUF || nd_in > P.dcap[s]) { failed = true; break; }
    for (int idx = tid; idx < S * S; idx += SF_NT) {
      const int j = idx / S, i = idx - j * S;
      if (i >= j) F[i + j * SF_LDF] = 0.0;
    }
    for (int i = tid; i < nc; i += SF_NT) fid[i] = P.cols[c0 + i];
    for (int i = tid; i < ncb; i += SF_NT) fid[fs + i] = P.rows[r0 + i];
    __syncthreads();
    // ---- original entries (unique targets; sources summed in input order) ----
    for (int e = P.ent_ptr[s] + tid; e < P.ent_ptr[s + 1]; e += SF_NT) {
      double v = 0.0;
      for (int p = P.tgt_src_ptr[e]; p < P.tgt_src_ptr[e + 1]; ++p) v += vals[B.val_off + P.tgt_src[p]];
      int r = P.tgt_row[e];
      const int c = P.tgt_col[e];
      if (r >= nc) r += nd_in;
      F[r + c * SF_LDF] = v;
    }
    __syncthreads();
    // ---- extend-add the children (top of stack first) ----
    int off = 0;
    for (int c = 0; c < nch; ++c) {
      const int *ft = reinterpret_cast<const int *>(B.stack + sp - 2);
      const int dim = ft[0], ndo = ft[1], child = ft[2];
      const long long base = sp - rec_size(dim);
      const int *ids = reinterpret_cast<const int *>(B.stack + base);
      const double *M = B.stack + base + (dim + 1) / 2;
      const int *crel = P.rel + P.row_ptr[child];
      for (int i = tid; i < dim; i += SF_NT) {
        if (i < ndo) { map[i] = nc + off + i; fid[nc + off + i] = ids[i]; }
        else { const int rr = crel[i - ndo]; map[i] = rr < nc ? rr : rr + nd_in; }
      }
      __syncthreads();
      for (int idx = tid; idx < dim * dim; idx += SF_NT) {
        const int j = idx / dim, i = idx - j * dim;
        if (i < j) continue;
        const int a = map[i], b = map[j];
        fent(F, a, b) += M[i + (long long)j * dim];
      }
      __syncthreads();
      off += ndo;
      sp = base;
    }
    // ---- partial factorisation ----
    const int ne = factor_front(F, S, fs, fid, bsz, w0, w1, u, pivtol, cnt, sh_i);
    const int ndo = fs - ne, dim = S - ne;
    // ---- store L / D, ids, pivot flags ----
    {
      const int caprows = nc + P.dcap[s] + ncb;
      double *Ls = B.L + P.l_off[s];
      for (int idx = tid; idx < S * ne; idx += SF_NT) {
        const int j = idx / S, i = idx - j * S;
        if (i >= j) Ls[i + (long long)j * caprows] = F[i + j * SF_LDF];
      }
      int *fo = B.fid + P.fid_off[s];
      for (int i = tid; i < S; i += SF_NT) fo[i] = fid[i];
      int *po = B.pbz + P.fs_off[s];
      for (int i = tid; i < ne; i += SF_NT) po[i] = bsz[i];
      if (tid == 0) { B.meta[2 * s] = ne; B.meta[2 * s + 1] = S; }
    }
    // ---- contribution block: to the parent's stack record, or into the dense root front ----
    if (P.parent[s] >= 0) {
      const long long need = rec_size(dim);
      if (sp + need > B.stack_cap) { failed = true; break; }
      int *ids = reinterpret_cast<int *>(B.stack + sp);
      double *M = B.stack + sp + (dim + 1) / 2;
      for (int i = tid; i < ndo; i += SF_NT) ids[i] = fid[ne + i];
      for (int idx = tid; idx < dim * dim; idx += SF_NT) {
        const int j = idx / dim, i = idx - j * dim;
        if (i >= j) M[i + (long long)j * dim] = F[(ne + i) + (ne + j) * SF_LDF];
      }
      if (tid == 0) {
        int *ft = reinterpret_cast<int *>(B.stack + sp + need - 2);
        ft[0] = dim; ft[1] = ndo; ft[2] = s; ft[3] = 0;
      }
      sp += need;
    } else {
      if (ndroot + ndo > P.DR) { failed = true; break; }
      const int *crel = P.rel + r0;
      for (int i = tid; i < dim; i += SF_NT) {
        if (i < ndo) { map[i] = P.nT + ndroot + i; B.rootids[P.nT + ndroot + i] = fid[ne + i]; }
        else map[i] = crel[i - ndo];
      }
      __syncthreads();
      for (int idx = tid; idx < dim * dim; idx += SF_NT) {
        const int j = idx / dim, i = idx - j * dim;
        if (i < j) continue;
        int a = map[i], b = map[j];
        if (a < b) { const int x = a; a = b; b = x; }
        R.A[(size_t)a + (size_t)b * R.ld] += F[(ne + i) + (ne + j) * SF_LDF];
      }
      ndroot += ndo;
    }
    __syncthreads();
  }
  // ---- finish the root: identity on unused delayed slots, ids of the static columns ----
  for (int i = tid; i < P.nT; i += SF_NT) B.rootids[i] = P.rootcols[i];
  for (int t = ndroot + tid; t < P.DR; t += SF_NT) {
    B.rootids[P.nT + t] = -1;
    R.A[(size_t)(P.nT + t) + (size_t)(P.nT + t) * R.ld] = 1.0;
  }
  __syncthreads();
  if (tid == 0) {
    B.info[0] = failed ? 1 : 0;
    B.info[1] = ndroot;
    // the DR - ndroot identity slots will be counted as positive pivots by the root: cancel them here
    const long long pos = (long long)cnt[0] - (long long)(P.DR - ndroot);
    atomicAdd(&inertia[0], (unsigned long long)pos);
    atomicAdd(&inertia[1], (unsigned long long)cnt[1]);
    atomicAdd(&inertia[2], (unsigned long long)cnt[2]);
  }
}

// ---- solves on the subtree part ---------------------------------------------------------------
// forward: y <- rhs; for every front (postorder) z = L11^-1 y[elim], y[rest] -= L21 z; then the root
// right-hand side is gathered.  y keeps z for the backward sweep.
__global__ void __launch_bounds__(SF_NT) subtree_forward_kernel(const SparseBlock *__restrict__ blocks,
                                                                const PlanDev *__restrict__ plans,
                                                                const double *__restrict__ rhs,
                                                                const long long *__restrict__ vec_off,
                                                                double *__restrict__ ywork,
                                                                double *__restrict__ root_rhs,
                                                                const long long *__restrict__ root_off) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  double *Ls = reinterpret_cast<double *>(sm_raw);  // S x ne block, pitch SF_LDF
  double *z = Ls + SF_SBUF * SF_LDF;
  int *fid = reinterpret_cast<int *>(z + SF_SBUF);
  int *bsz = fid + SF_SBUF;
  const SparseBlock B = blocks[blockIdx.x];
  const PlanDev P = plans[B.plan];
  const int tid = threadIdx.x, lane = tid & 31;
  const double *r = rhs + vec_off[blockIdx.x];
  double *y = ywork + vec_off[blockIdx.x];
  for (int i = tid; i < P.n; i += SF_NT) y[i] = r[i];
  __syncthreads();
  for (int s = 0; s < P.ns; ++s) {
    const int ne = B.meta[2 * s], S = B.meta[2 * s + 1];
    if (ne == 0) continue;
    const int nc = P.col_ptr[s + 1] - P.col_ptr[s], ncb = P.row_ptr[s + 1] - P.row_ptr[s];
    const int caprows = nc + P.dcap[s] + ncb;
    const double *Lg = B.L + P.l_off[s];
    for (int idx = tid; idx < S * ne; idx += SF_NT) {
      const int j = idx / S, i = idx - j * S;
      if (i > j) Ls[i + j * SF_LDF] = Lg[i + (long long)j * caprows];
    }
    for (int i = tid; i < S; i += SF_NT) fid[i] = B.fid[P.fid_off[s] + i];
    for (int i = tid; i < ne; i += SF_NT) bsz[i] = B.pbz[P.fs_off[s] + i];
    __syncthreads();
    for (int i = tid; i < ne; i += SF_NT) z[i] = y[fid[i]];
    __syncthreads();
    if (tid < 32) {
      for (int c = 0; c < ne; ++c) {
        const double zc = z[c];
        const int skip = bsz[c] == 2 ? c + 1 : -1;
        for (int i = c + 1 + lane; i < ne; i += 32)
          if (i != skip) z[i] -= Ls[i + c * SF_LDF] * zc;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = ne + tid; i < S; i += SF_NT) {
      double acc = 0.0;
      for (int c = 0; c < ne; ++c) acc += Ls[i + c * SF_LDF] * z[c];
      y[fid[i]] -= acc;
    }
    for (int i = tid; i < ne; i += SF_NT) y[fid[i]] = z[i];
    __syncthreads();
  }
  double *rr = root_rhs + root_off[blockIdx.x];
  for (int p = tid; p < P.nT + P.DR; p += SF_NT) {
    const int id = B.rootids[p];
    rr[p] = id >= 0 ? y[id] : 0.0;
  }
}

// backward: x[root ids] <- root solution; for every front in reverse postorder
//   x[elim] = L11^-T (D^-1 z - L21^T x[rest]).
__global__ void __launch_bounds__(SF_NT) subtree_backward_kernel(const SparseBlock *__restrict__ blocks,
                                                                 const PlanDev *__restrict__ plans,
                                                                 const double *__restrict__ ywork,
                                                                 const long long *__restrict__ vec_off,
                                                                 const double *__restrict__ root_x,
                                                                 const long long *__restrict__ root_off,
                                                                 double *__restrict__ xout) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  double *Ls = reinterpret_cast<double *>(sm_raw);
  double *z = Ls + SF_SBUF * SF_LDF;
  double *xr = z + SF_SBUF;
  int *fid = reinterpret_cast<int *>(xr + SF_SBUF);
  int *bsz = fid + SF_SBUF;
  const SparseBlock B = blocks[blockIdx.x];
  const PlanDev P = plans[B.plan];
  const int tid = threadIdx.x, lane = tid & 31;
  const double *y = ywork + vec_off[blockIdx.x];
  double *x = xout + vec_off[blockIdx.x];
  const double *rx = root_x + root_off[blockIdx.x];
  for (int p = tid; p < P.nT + P.DR; p += SF_NT) {
    const int id = B.rootids[p];
    if (id >= 0) x[id] = rx[p];
  }
  __syncthreads();
  for (int s = P.ns - 1; s >= 0; --s) {
    const int ne = B.meta[2 * s], S = B.meta[2 * s + 1];
    if (ne == 0) continue;
    const int nc = P.col_ptr[s + 1] - P.col_ptr[s], ncb = P.row_ptr[s + 1] - P.row_ptr[s];
    const int caprows = nc + P.dcap[s] + ncb;
    const double *Lg = B.L + P.l_off[s];
    for (int idx = tid; idx < S * ne; idx += SF_NT) {
      const int j = idx / S, i = idx - j * S;
      if (i >= j) Ls[i + j * SF_LDF] = Lg[i + (long long)j * caprows];
    }
    for (int i = tid; i < S; i += SF_NT) fid[i] = B.fid[P.fid_off[s] + i];
    for (int i = tid; i < ne; i += SF_NT) bsz[i] = B.pbz[P.fs_off[s] + i];
    __syncthreads();
    for (int i = ne + tid; i < S; i += SF_NT) xr[i] = x[fid[i]];
    // w = D^-1 z
    for (int k = tid; k < ne; k += SF_NT) {
      const int b = bsz[k];
      if (b == 1) {
        const double d = Ls[k + k * SF_LDF];
        z[k] = d != 0.0 ? y[fid[k]] / d : 0.0;
      } else if (b == 2) {
        const double e21 = Ls[k + 1 + k * SF_LDF];
        const double akm1 = Ls[k + k * SF_LDF] / e21, ak = Ls[k + 1 + (k + 1) * SF_LDF] / e21;
        const double denom = akm1 * ak - 1.0;
        const double bkm1 = y[fid[k]] / e21, bk = y[fid[k + 1]] / e21;
        z[k] = (ak * bkm1 - bk) / denom;
        z[k + 1] = (akm1 * bk - bkm1) / denom;
      }
    }
    __syncthreads();
    // z[c] -= L21(:,c)^T xr
    for (int c = tid; c < ne; c += SF_NT) {
      double acc = 0.0;
      for (int i = ne; i < S; ++i) acc += Ls[i + c * SF_LDF] * xr[i];
      z[c] -= acc;
    }
    __syncthreads();
    if (tid < 32) {
      for (int rr = ne - 1; rr > 0; --rr) {
        const double xv = z[rr];
        for (int c = lane; c < rr; c += 32)
          if (!(bsz[c] == 2 && rr == c + 1)) z[c] -= Ls[rr + c * SF_LDF] * xv;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = tid; i < ne; i += SF_NT) x[fid[i]] = z[i];
    __syncthreads();
  }
}

// worst failure flag over the sparse blocks -> flag[1]
__global__ void collect_sparse_info_kernel(const SparseBlock *__restrict__ blocks, int count, int *flag) {
  int bad = 0;
  for (int b = threadIdx.x; b < count; b += blockDim.x)
    if (blocks[b].info[0] != 0) bad = 1;
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flag[1] = bad;
}

}  // namespace ppb
