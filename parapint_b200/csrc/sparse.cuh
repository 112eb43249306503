// Multifrontal factorisation / solves of the SUBTREE part of every diagonal block (one CTA per block).
//
// The block's assembly tree is processed level by level (level 0 = leaves).  Fronts of one level
// are independent: tiny ones are taken by single warps concurrently (leaves even by parts of a warp, GPU-wide),
// medium ones by two-warp groups eight at a time, larger ones by the whole CTA.
// A front is assembled in shared memory (original entries + the contribution blocks of its
// children, read from per-front slots in HBM), partially factorised with threshold pivoting among
// its fully-summed rows (1x1 and 2x2 pivots; columns that find no acceptable pivot are DELAYED to the
// parent, MA57-style), its L/D columns go to the block's factor arena and its contribution block to
// its slot.  Children of the root are finally added into the block's dense root front, which
// front_small_kernel (small.cuh) or the batched Bunch-Kaufman kernels of factor.cuh finish (they also see the
// delayed columns, so no pivot is ever forced: the inertia stays exact).  All sums run in a fixed order: results are reproducible.
#pragma once
#include <cooperative_groups.h>
#include "front.cuh"

namespace ppb {

#ifdef PP_TRACE
// debugging aid (tools/trace_levels.py): cycle stamps of block 0, thread 0
__device__ long long g_trace[2048];
#define PP_TR(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (slot) < 2048) g_trace[(slot)] = clock64(); } while (0)
__device__ int g_tp;
// phase stamps inside process_front / forward_front (group 0 of block 0 only): (cycles << 4) | phase
#define PP_TRP(ph) do { if (G == SF_MG && blockIdx.x == 0 && threadIdx.x == 0) { const int q_ = g_tp++; if (q_ < 760) g_trace[256 + q_] = (clock64() << 4) | (ph); } } while (0)
#else
#define PP_TR(slot) do { } while (0)
#define PP_TRP(ph) do { } while (0)
#endif

// Static description of one supernode, fetched with a single 80-byte load.
struct __align__(16) SnHead {
  int c0, nc, r0, ncb;            // own columns: cols[c0..c0+nc); contribution rows: rows[r0..r0+ncb)
  int ch0, nch, dcap, dslot;      // children: child_idx[ch0..ch0+nch); delayed-in / delayed-out capacity
  int fid_off, fs_off, vec_off, ent0;
  int nent, parent, pad0, pad1;   // original-entry targets: tgt[ent0..ent0+nent)
  long long l_off, cb_off;
};

struct PlanDev {
  int n, m, nT, DR, ns, nlevels, nrootch, pad;
  const SnHead *heads;
  const int *rootcols, *cols, *rows, *rel, *child_idx, *root_children;
  const int *tiny_ptr, *tiny_idx, *med_ptr, *med_idx, *big_ptr, *big_idx;
  const int2 *tgt;      // .x = row | col << 8 | count << 16,  .y = source index (count == 1) or offset into tgt_src
  const int *tgt_src;
};

struct SparseBlock {
  int plan, root;        // plan index; index of the dense root front
  long long val_off;     // base index of this block's input values
  long long slot_base;   // index of shift slot 0 (behind the input values): a source s < 0 reads slot -s - 1
  double *L;             // factor arena of the subtree supernodes
  double *cb;            // contribution slots
  double *vec;           // forward-solve contribution vectors
  int *fid;              // per supernode: original ids of the front rows after pivoting
  int *pbz;              // per supernode: pivot flags (1, 2, 0) of the eliminated columns
  int *opos;             // per supernode: pre-pivoting position of every fully-summed row
  int *meta;             // per supernode: {ne, S, nd_out}
  int *rootids;          // [nT + DR] original id per root position, -1 for unused delayed slots
  int *info;             // [0] failure flag, [1] delayed pivots that reached the root
};

constexpr int SF_NT = 512;           // threads per CTA in the subtree kernels
constexpr int SF_NW = SF_NT / 32;
constexpr int SF_SBUF = 96;          // largest front held in shared memory (whole CTA)
constexpr int SF_LDF = SF_SBUF + 1;  // odd pitch: conflict-free row and column walks
constexpr int SF_TBUF = 24;          // largest front held by a single warp
constexpr int SF_TLD = SF_TBUF + 1;

// shared-memory work area of one group (a warp or the whole CTA)
struct FrontBuf {
  double *F;   // cap x ld, lower triangle
  double *w0, *w1;
  int *fid, *opos, *bsz, *map;
  int *sh;     // 8 ints of scratch
  int ld, cap;
};

__host__ __device__ constexpr size_t fb_bytes(int cap, int ld) {
  return (size_t)cap * ld * 8 + 2 * (size_t)cap * 8 + 4 * (size_t)cap * 4 + 32;
}
constexpr size_t SF_BIG_BYTES = fb_bytes(SF_SBUF, SF_LDF);
constexpr size_t SF_TINY_BYTES = (fb_bytes(SF_TBUF, SF_TLD) + 15) / 16 * 16;
// staging area used when a front has many small children: their contribution entries are fetched
// by all threads at once (one child per thread) and then applied in child order by a single warp
constexpr int SF_STG = 2048;     // staged entries
constexpr int SF_MAXCH = 512;    // children per staging batch
constexpr int SF_CHDIM = 12;     // largest child contribution block that is staged
// two-warp groups: fronts of up to SF_MBUF rows, each group with its own staging area
constexpr int SF_MBUF = 32, SF_MLD = SF_MBUF + 1, SF_MSTG = 1024, SF_MMAXCH = 256;
constexpr int SF_MG = 64, SF_NG = 512 / SF_MG;  // medium fronts: eight groups of two warps (named barriers 1..8)


struct Stage {
  double *val;   // [cap] staged values
  int *tgt;      // [cap] offset into F (or position in v)
  int *cnt;      // [maxch + 1] entry offsets per child
  int *ndo;      // [maxch] delayed columns per child / running offsets
  int *big;      // [maxch] children handled one by one afterwards
  int cap, maxch;
};

__host__ __device__ constexpr size_t stage_bytes(int cap, int maxch) {
  return ((size_t)cap * 12 + (size_t)(3 * maxch + 16) * 4 + 15) / 16 * 16;
}
__host__ __device__ constexpr size_t align16(size_t v) { return (v + 15) / 16 * 16; }
__host__ __device__ constexpr size_t max3(size_t a, size_t b, size_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
constexpr size_t SF_BIG_FB = align16(fb_bytes(SF_SBUF, SF_LDF));
constexpr size_t SF_MED_FB = align16(fb_bytes(SF_MBUF, SF_MLD));
constexpr size_t SF_MED_BYTES = SF_MED_FB + stage_bytes(SF_MSTG, SF_MMAXCH);
constexpr size_t SF_WORK = max3(SF_BIG_FB + stage_bytes(SF_STG, SF_MAXCH), SF_NG * SF_MED_BYTES, SF_NW * SF_TINY_BYTES);
constexpr size_t SF_SMEM = SF_WORK + 96 * sizeof(int);

__device__ __forceinline__ Stage carve_stage(unsigned char *base, int cap = SF_STG, int maxch = SF_MAXCH) {
  Stage g;
  g.val = reinterpret_cast<double *>(base);
  g.tgt = reinterpret_cast<int *>(g.val + cap);
  g.cnt = g.tgt + cap;
  g.ndo = g.cnt + maxch + 8;
  g.big = g.ndo + maxch;
  g.cap = cap;
  g.maxch = maxch;
  return g;
}

__device__ __forceinline__ FrontBuf carve(unsigned char *base, int cap, int ld) {
  FrontBuf b;
  b.F = reinterpret_cast<double *>(base);
  b.w0 = b.F + (size_t)cap * ld;
  b.w1 = b.w0 + cap;
  b.fid = reinterpret_cast<int *>(b.w1 + cap);
  b.opos = b.fid + cap;
  b.bsz = b.opos + cap;
  b.map = b.bsz + cap;
  b.sh = b.map + cap;
  b.ld = ld;
  b.cap = cap;
  return b;
}

// groups: part of a warp (8 or 16 lanes: several tiny leaf fronts share a warp), a warp (32), two or four warps
// (64 / 128 threads, named barrier 1 + group index) or the whole CTA
template <int G> struct Grp {
  static constexpr int LW = G < 32 ? G : 32;       // lanes that walk a column together
  static constexpr int NW = G < 32 ? 1 : G / 32;   // warps of the group
};
template <int G> __device__ __forceinline__ unsigned gmask() {  // lanes of this thread's group within its warp
  return G < 32 ? (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1))) : 0xffffffffu;
}
template <int G> __device__ __forceinline__ int gtid() {
  return G <= 128 ? (int)(threadIdx.x & (G - 1)) : (int)threadIdx.x;
}
template <int G> __device__ __forceinline__ void gsync() {
  if (G <= 32) __syncwarp(gmask<G>());
  else if (G == 128) asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory");
  else if (G == 64) asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)(threadIdx.x >> 6)) : "memory");
  else __syncthreads();
}

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

__device__ __forceinline__ double &fent(double *F, int ld, int r, int c) {  // lower-triangle accessor
  return r >= c ? F[r + c * ld] : F[c + r * ld];
}

// symmetric interchange of positions a < b in the lower-stored front (rows of L included)
template <int G>
__device__ __forceinline__ void front_swap(const FrontBuf &B, int S, int a, int b) {
  if (a == b) return;
  double *F = B.F;
  const int ld = B.ld;
  for (int i = gtid<G>(); i < S; i += G) {
    double *p, *q;
    if (i < a) { p = &F[a + i * ld]; q = &F[b + i * ld]; }
    else if (i == a) { p = &F[a + a * ld]; q = &F[b + b * ld]; }
    else if (i < b) { p = &F[i + a * ld]; q = &F[b + i * ld]; }
    else if (i == b) continue;
    else { p = &F[i + a * ld]; q = &F[i + b * ld]; }
    const double t = *p;
    *p = *q;
    *q = t;
  }
  if (gtid<G>() == 0) {
    int t = B.fid[a]; B.fid[a] = B.fid[b]; B.fid[b] = t;
    t = B.opos[a]; B.opos[a] = B.opos[b]; B.opos[b] = t;
  }
  gsync<G>();
}

// Partial factorisation of the S x S front; the first fs rows are fully summed.  Returns the number
// of eliminated columns; inertia counts go to cnt[3] (shared, atomics on integers).
template <int G>
__device__ int factor_front(const FrontBuf &B, int S, int fs, double u, double pivtol, int *cnt, int ntest = -1) {
  if (ntest < 0) ntest = S;  // rows that take part in the threshold tests (a dense root excludes its border rows)
  constexpr int LW = Grp<G>::LW, NW = Grp<G>::NW;
  const int tid = gtid<G>(), lane = tid & (LW - 1), gw = G < 32 ? 0 : tid >> 5;
  const unsigned wmask = gmask<G>();
  double *F = B.F;
  const int ld = B.ld;
  int t = 0;
  int npos = 0, nneg = 0, nzero = 0;  // inertia of the 1x1 fast paths, kept by thread 0 and flushed once
  while (t < fs) {
    if (NW == 1 && S <= LW) {
      // One row per lane (front no taller than the group): the pivot column lives in a register, its entries reach
      // the other lanes by shuffle, and every lane updates only its own row -- one warp-level barrier per column.
      // Same test and same arithmetic as the general 1x1 step below.
      const bool mine = lane > t && lane < S;
      const double d = F[t + t * ld];
      const double rd = 1.0 / d;  // issued ahead of the column reduction it does not depend on
      const double w = mine ? F[lane + t * ld] : 0.0;
      unsigned kmax = (mine && lane < ntest) ? ((unsigned)__double2hiint(w) & 0x7fffffffu) : 0u;
      kmax = __reduce_max_sync(wmask, kmax);
      const double cmax = kmax ? __hiloint2double((int)kmax, -1) : 0.0;
      if (fabs(d) > pivtol && fabs(d) >= u * cmax) {
        const double ws = w * rd;
        // eight columns per batch: shuffles and loads of a batch are issued together, then the FMAs, then the
        // stores (a column-at-a-time loop would serialise one shuffle + one shared-memory round trip per column)
        for (int j0 = t + 1; j0 < S; j0 += 8) {
          double wj[8], f[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            wj[q] = __shfl_sync(wmask, ws, j < S ? j : t, LW);
            f[q] = (j < S && lane >= j && lane < S) ? F[lane + j * ld] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            if (j < S && lane >= j && lane < S) F[lane + j * ld] = f[q] - w * wj[q];
          }
        }
        if (mine) F[lane + t * ld] = ws;
        if (tid == 0) {
          B.bsz[t] = 1;
          if (d > 0.0) ++npos; else if (d < 0.0) ++nneg; else ++nzero;
        }
        gsync<G>();
        t += 1;
        continue;
      }
    } else if (LW == 32 && S <= 64) {
      // Fronts of up to 64 rows on several warps (the 50 x 50 pivot blocks of the config-2 roots and coupling front, the
      // whole-CTA fronts of the subtree): every warp keeps the pivot column in registers, TWO rows per lane (rows
      // `lane` and `lane + 32`), decides the test redundantly as below, and updates its share of the columns with the
      // multipliers taken from those registers by shuffle -- no dependent shared-memory load per column, the loads
      // of a batch of four columns issued together.  Same test, same arithmetic, same results as the path below.
      // (Measured at config 2: small fronts 0.113 -> 0.108 ms, subtree 0.178 -> 0.175 ms per step; the roots' pivot blocks
      // are multiplier columns with zero diagonals and mostly take the general search below.)
      const int r0 = lane, r1 = lane + 32;
      const bool in0 = r0 > t && r0 < S, in1 = r1 > t && r1 < S;
      const double d = F[t + t * ld];
      const double c0 = in0 ? F[r0 + t * ld] : 0.0, c1 = in1 ? F[r1 + t * ld] : 0.0;
      const double rd = 1.0 / d;
      unsigned kmax = max((in0 && r0 < ntest) ? ((unsigned)__double2hiint(c0) & 0x7fffffffu) : 0u,
                          (in1 && r1 < ntest) ? ((unsigned)__double2hiint(c1) & 0x7fffffffu) : 0u);
      kmax = __reduce_max_sync(wmask, kmax);
      const double cmax = kmax ? __hiloint2double((int)kmax, -1) : 0.0;
      if (fabs(d) > pivtol && fabs(d) >= u * cmax) {
        for (int j0 = t + 1 + gw; j0 < S; j0 += 4 * NW) {
          double wj[4], f0[4], f1[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q * NW, js = j < S ? j : t + 1;   // (uniform over the warp)
            wj[q] = __shfl_sync(wmask, js < 32 ? c0 : c1, js & 31) * rd;
            f0[q] = (j < S && r0 >= j) ? F[r0 + j * ld] : 0.0;
            f1[q] = (j < S && r1 >= j && r1 < S) ? F[r1 + j * ld] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q * NW;
            if (j < S && r0 >= j) F[r0 + j * ld] = f0[q] - c0 * wj[q];
            if (j < S && r1 >= j && r1 < S) F[r1 + j * ld] = f1[q] - c1 * wj[q];
          }
        }
        if (tid == 0) {
          B.bsz[t] = 1;
          if (d > 0.0) ++npos; else if (d < 0.0) ++nneg; else ++nzero;
        }
        gsync<G>();
        if (gw == 0) {   // nobody reads column t any more
          if (in0) F[r0 + t * ld] = c0 * rd;
          if (in1) F[r1 + t * ld] = c1 * rd;
        }
        t += 1;
        continue;
      }
    } else {
      // Fast path, decided redundantly (and identically) by every warp, so nothing has to be published: the
      // diagonal of the leading column passes the threshold test as it is -- the first candidate the general
      // search below would accept.  One barrier per eliminated column instead of four.
      const double d = F[t + t * ld];
      const double rd = 1.0 / d;  // issued ahead of the column scan it does not depend on
      unsigned kmax = 0u;
      for (int i = t + 1 + lane; i < ntest; i += LW)
        kmax = max(kmax, (unsigned)__double2hiint(F[i + t * ld]) & 0x7fffffffu);
      kmax = __reduce_max_sync(wmask, kmax);
      const double cmax = kmax ? __hiloint2double((int)kmax, -1) : 0.0;
      if (fabs(d) > pivtol && fabs(d) >= u * cmax) {
        for (int j = t + 1 + gw; j < S; j += NW) {
          const double wj = F[j + t * ld] * rd;
          for (int i = j + lane; i < S; i += LW) F[i + j * ld] -= F[i + t * ld] * wj;
        }
        if (tid == 0) {
          B.bsz[t] = 1;
          if (d > 0.0) ++npos; else if (d < 0.0) ++nneg; else ++nzero;
        }
        gsync<G>();
        for (int i = t + 1 + tid; i < S; i += G) F[i + t * ld] *= rd;  // nobody reads column t any more
        t += 1;
        continue;
      }
    }
    if (gw == 0) {
      // Pivot search.  Column maxima only feed threshold tests, so they are reduced on the high 32
      // bits of |value| (monotone, one redux.sync each) and widened to an upper bound; the partner
      // row for a 2x2 pivot is the (near-)largest fully-summed entry, lowest index on ties.
      int kind = 0, pc = -1, pr = -1;
      for (int c = t; c < fs; ++c) {
        unsigned kmax = 0u, kbest = 0u;
        for (int i = t + lane; i < ntest; i += LW) {
          if (i == c) continue;
          const unsigned hi = (unsigned)__double2hiint(fent(F, ld, i, c)) & 0x7fffffffu;
          kmax = max(kmax, hi);
          if (i < fs) kbest = max(kbest, (hi & 0xffffff00u) | (unsigned)(255 - i));
        }
        kmax = __reduce_max_sync(wmask, kmax);
        kbest = __reduce_max_sync(wmask, kbest);
        const double cmax = kmax ? __hiloint2double((int)kmax, -1) : 0.0;
        const double dcc = F[c + c * ld];
        if (fabs(dcc) > pivtol && fabs(dcc) >= u * cmax) { kind = 1; pc = c; break; }
        if (kbest != 0u) {
          const int r = 255 - (int)(kbest & 255u);  // fronts have at most 256 rows
          const double b = fent(F, ld, r, c);
          if (fabs(b) > pivtol) {
            unsigned kc = 0u, kr = 0u;
            for (int i = t + lane; i < ntest; i += LW) {
              if (i == c || i == r) continue;
              kc = max(kc, (unsigned)__double2hiint(fent(F, ld, i, c)) & 0x7fffffffu);
              kr = max(kr, (unsigned)__double2hiint(fent(F, ld, i, r)) & 0x7fffffffu);
            }
            kc = __reduce_max_sync(wmask, kc);
            kr = __reduce_max_sync(wmask, kr);
            const double cm_c = kc ? __hiloint2double((int)kc, -1) : 0.0;
            const double cm_r = kr ? __hiloint2double((int)kr, -1) : 0.0;
            const double drr = F[r + r * ld];
            const double det = dcc * drr - b * b;
            const double g1 = (fabs(drr) * cm_c + fabs(b) * cm_r) / fabs(det);
            const double g2 = (fabs(b) * cm_c + fabs(dcc) * cm_r) / fabs(det);
            if (g1 * u <= 1.0 && g2 * u <= 1.0) { kind = 2; pc = c; pr = r; break; }
            if (fabs(drr) > pivtol && fabs(drr) >= u * fmax(cm_r, fabs(b))) { kind = 1; pc = r; break; }
          }
        }
      }
      if (lane == 0) { B.sh[0] = kind; B.sh[1] = pc; B.sh[2] = pr; }
#ifdef PP_TRACE
      if (G == SF_MG && blockIdx.x == 0 && threadIdx.x == 0) {
        g_trace[2040 + kind] += 1;              // steps by kind (0 = no pivot, 1, 2)
        g_trace[2043] += (pc >= 0 ? (kind == 2 ? min(pc, pr) : pc) - t + 1 : fs - t);  // candidates examined (approx.)
        g_trace[2044] = clock64();
      }
#endif
    }
    gsync<G>();
    const int kind = B.sh[0];
    int pc = B.sh[1], pr = B.sh[2];
    gsync<G>();
    if (kind == 0) break;
    if (kind == 2 && pr < pc) { const int x = pc; pc = pr; pr = x; }
    front_swap<G>(B, S, t, pc);
    if (kind == 1) {
      const double d = F[t + t * ld];
      for (int i = t + 1 + tid; i < S; i += G) B.w0[i] = F[i + t * ld];
      gsync<G>();
      const double rd = 1.0 / d;
      for (int j = t + 1 + gw; j < S; j += NW) {
        const double wj = B.w0[j] * rd;
        for (int i = j + lane; i < S; i += LW) F[i + j * ld] -= B.w0[i] * wj;
      }
      for (int i = t + 1 + tid; i < S; i += G) F[i + t * ld] = B.w0[i] * rd;
      if (tid == 0) {
        B.bsz[t] = 1;
        atomicAdd(&cnt[d > 0.0 ? 0 : (d < 0.0 ? 1 : 2)], 1);
      }
      gsync<G>();
      t += 1;
    } else {
      front_swap<G>(B, S, t + 1, pr);
      const double e11 = F[t + t * ld], e21 = F[t + 1 + t * ld], e22 = F[t + 1 + (t + 1) * ld];
      for (int i = t + 2 + tid; i < S; i += G) {
        B.w0[i] = F[i + t * ld];
        B.w1[i] = F[i + (t + 1) * ld];
      }
      gsync<G>();
      const double d11 = e22 / e21, d22 = e11 / e21;
      const double sc = (1.0 / (d11 * d22 - 1.0)) / e21;
      for (int j = t + 2 + gw; j < S; j += NW) {
        const double a0 = B.w0[j], a1 = B.w1[j];
        for (int i = j + lane; i < S; i += LW) {
          const double l0 = sc * (d11 * B.w0[i] - B.w1[i]), l1 = sc * (d22 * B.w1[i] - B.w0[i]);
          F[i + j * ld] -= l0 * a0 + l1 * a1;
        }
      }
      for (int i = t + 2 + tid; i < S; i += G) {
        F[i + t * ld] = sc * (d11 * B.w0[i] - B.w1[i]);
        F[i + (t + 1) * ld] = sc * (d22 * B.w1[i] - B.w0[i]);
      }
      if (tid == 0) {
        B.bsz[t] = 2;
        B.bsz[t + 1] = 0;
        const double det = d11 * d22 - 1.0;  // sign of the determinant (scaled by e21^2 > 0)
        if (det < 0.0) { atomicAdd(&cnt[0], 1); atomicAdd(&cnt[1], 1); }
        else if (det > 0.0) atomicAdd(&cnt[e11 + e22 > 0.0 ? 0 : 1], 2);
        else { atomicAdd(&cnt[2], 1); atomicAdd(&cnt[e11 + e22 > 0.0 ? 0 : 1], 1); }
      }
      gsync<G>();
      t += 2;
    }
  }
  if (tid == 0) {
    if (npos) atomicAdd(&cnt[0], npos);
    if (nneg) atomicAdd(&cnt[1], nneg);
    if (nzero) atomicAdd(&cnt[2], nzero);
  }
  return t;
}

// Offsets of a front's children in the staging area, computed by the first warp of the group: ndo[k] becomes the
// exclusive prefix of the delayed-column counts, cnt[k] the exclusive prefix of the staged-entry counts
// (cnt[nch] = total); children with a negative count are listed in big[] (index order).  Returns through
// sh[4] = delayed columns in, sh[5] = number of big children, sh[7] = 1 if the staged entries exceed the
// capacity (the caller then falls back to a serial first-fit pass).
__device__ __forceinline__ void stage_prefix_warp(const Stage &stg, int nch, int *sh) {
  const int lane = threadIdx.x & 31;
  int offd = 0, ent = 0, nbig = 0;
  for (int base = 0; base < nch; base += 32) {
    const int k = base + lane;
    const bool valid = k < nch;
    const int d = valid ? stg.ndo[k] : 0, e = valid ? stg.cnt[k] : 0;
    const bool isbig = valid && e < 0;
    const int ep = e > 0 ? e : 0;
    int ds = d, es = ep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t0 = __shfl_up_sync(0xffffffffu, ds, o), t1 = __shfl_up_sync(0xffffffffu, es, o);
      if (lane >= o) { ds += t0; es += t1; }
    }
    const unsigned bm = __ballot_sync(0xffffffffu, isbig);
    if (valid) { stg.ndo[k] = offd + ds - d; stg.cnt[k] = ent + es - ep; }
    if (isbig) stg.big[nbig + __popc(bm & ((1u << lane) - 1u))] = k;
    offd += __shfl_sync(0xffffffffu, ds, 31);
    ent += __shfl_sync(0xffffffffu, es, 31);
    nbig += __popc(bm);
  }
  if (lane == 0) { stg.cnt[nch] = ent; sh[4] = offd; sh[5] = nbig; sh[7] = ent > stg.cap ? 1 : 0; }
}

// dst[tgt[e]] += val[e] over the staged entries e = 0..total-1, reproducibly: the entries are cut into rounds of 32;
// inside a round, lanes that hit the same target are found with match.any and their values are summed in lane
// (= staged = child) order; the rounds' partial sums are committed to dst in round order.  The warps of the group
// prepare NW rounds concurrently (that is the expensive part: match, shuffles) and then take turns committing, so a
// hub front with 85 tiny children costs a handful of barriers instead of 85 dependent shared-memory round trips.
template <int G>
__device__ __forceinline__ void apply_staged(double *dst, const int *tgt, const double *val, int total) {
  constexpr int NW = Grp<G>::NW;
  const int tid = gtid<G>(), lane = tid & 31, gw = tid >> 5;
  const int rounds = (total + 31) >> 5;
  for (int r0 = 0; r0 < rounds; r0 += NW) {
    const int r = r0 + gw;
    const int e = (r << 5) + lane;
    const bool on = r < rounds && e < total;
    const int t = on ? tgt[e] : -1 - lane;  // distinct dummies for idle lanes
    const double v = on ? val[e] : 0.0;
    const unsigned m = __match_any_sync(0xffffffffu, t);
    const int mx = __reduce_max_sync(0xffffffffu, __popc(m));
    double part = 0.0;
    unsigned rest = m;
    for (int it = 0; it < mx; ++it) {
      const int src = rest ? __ffs(rest) - 1 : lane;
      const double x = __shfl_sync(0xffffffffu, v, src);
      if (rest) part += x;
      rest &= rest - 1;
    }
    const bool leader = on && lane == __ffs(m) - 1;
    const int live = min(NW, rounds - r0);
    for (int q = 0; q < live; ++q) {
      if (gw == q && leader) dst[t] += part;
      if (NW > 1) gsync<G>(); else __syncwarp();
    }
  }
}

enum { PF_OK = 0, PF_DEFER = 1, PF_FAIL = 2 };

// Assemble, factor and store supernode s with group G.  PF_DEFER: the front (with its delayed
// columns) does not fit this group's buffer; nothing was written.
template <int G>
__device__ int process_front(const SparseBlock &Bk, const PlanDev &P, const double *__restrict__ vals, int s,
                             const FrontBuf &B, double u, double pivtol, int *cnt, const Stage &stg) {
  const int tid = gtid<G>();
  PP_TRP(0);
  const SnHead H = P.heads[s];
  const int nc = H.nc, ncb = H.ncb;
  const bool staged = G >= SF_MG && H.nch >= 4 && H.nch <= stg.maxch;
  const bool pref = G <= 32 && H.nch > 0 && H.nch <= Grp<G>::LW;
  int pc = 0, pcne = 0, pdim = 0, pndo = 0, pfid = 0, pr0 = 0;  // this lane's child (pref)
  long long pcb = 0;
  int nd_in = 0;
  if (staged) {
    // one child per thread: delayed count and staged-entry count; serial prefix by thread 0
    for (int k = tid; k < H.nch; k += G) {
      const int c = P.child_idx[H.ch0 + k];
      const int dim = Bk.meta[3 * c + 1] - Bk.meta[3 * c];
      stg.ndo[k] = Bk.meta[3 * c + 2];
      stg.cnt[k] = dim <= SF_CHDIM ? dim * (dim + 1) / 2 : -1;
    }
    gsync<G>();
    if (tid < 32) stage_prefix_warp(stg, H.nch, B.sh);
    gsync<G>();
    if (B.sh[7]) {  // staging area too small for all of them (rare): serial first-fit from the source counts
      if (tid == 0) {
        int off = 0, ent = 0, nbig = 0;
        for (int k = 0; k < H.nch; ++k) {
          const int c = P.child_idx[H.ch0 + k];
          const int dim = Bk.meta[3 * c + 1] - Bk.meta[3 * c];
          const int e = dim <= SF_CHDIM ? dim * (dim + 1) / 2 : -1;
          stg.ndo[k] = off;
          off += Bk.meta[3 * c + 2];
          if (e < 0 || ent + e > stg.cap) { stg.big[nbig++] = k; stg.cnt[k] = ent; }  // handled one by one
          else { stg.cnt[k] = ent; ent += e; }
          if (k + 1 == H.nch) stg.cnt[k + 1] = ent;
        }
        B.sh[4] = off;
        B.sh[5] = nbig;
      }
      gsync<G>();
    }
    nd_in = B.sh[4];
  } else if (pref) {
    // (sub-)warp groups: lane k fetches everything about child k at once -- two dependent round trips for all the
    // children instead of four per child; the loop below reads the fields by shuffle
    if (tid < H.nch) {
      pc = P.child_idx[H.ch0 + tid];
      pcne = Bk.meta[3 * pc];
      pdim = Bk.meta[3 * pc + 1] - pcne;
      pndo = Bk.meta[3 * pc + 2];
      const SnHead *C = P.heads + pc;
      pfid = C->fid_off;
      pr0 = C->r0;
      pcb = C->cb_off;
    }
    int v = tid < H.nch ? pndo : 0;
#pragma unroll
    for (int o = Grp<G>::LW / 2; o; o >>= 1) v += __shfl_xor_sync(gmask<G>(), v, o, Grp<G>::LW);
    nd_in = v;
  } else {
    for (int k = 0; k < H.nch; ++k) nd_in += Bk.meta[3 * P.child_idx[H.ch0 + k] + 2];
  }
  PP_TRP(1);
  const int fs = nc + nd_in, S = fs + ncb;
  if (S > B.cap) {
    if (G == SF_NT && tid == 0) { Bk.info[3] = 1; Bk.info[4] = s; Bk.info[5] = S; }
    return G == SF_NT ? PF_FAIL : PF_DEFER;
  }
  if (nd_in > H.dcap) {
    if (tid == 0) { Bk.info[3] = 2; Bk.info[4] = s; Bk.info[5] = nd_in; }
    return PF_FAIL;
  }
  double *F = B.F;
  const int ld = B.ld;
  constexpr int NWG = Grp<G>::NW, LWG = Grp<G>::LW;  // columns dealt over the warps of the group, rows over the lanes
  const int gwp = G < 32 ? 0 : tid >> 5, lanep = tid & (LWG - 1);
  for (int j = gwp; j < S; j += NWG)
    for (int i = j + lanep; i < S; i += LWG) F[i + j * ld] = 0.0;
  for (int i = tid; i < nc; i += G) B.fid[i] = P.cols[H.c0 + i];
  for (int i = tid; i < ncb; i += G) B.fid[fs + i] = P.rows[H.r0 + i];
  for (int i = tid; i < fs; i += G) B.opos[i] = i;
  gsync<G>();
  PP_TRP(2);
  // original entries: unique targets, sources summed in input order
  for (int e = tid; e < H.nent; e += G) {
    const int2 t = P.tgt[H.ent0 + e];
    const int cntv = t.x >> 16;
    double v;
    if (cntv == 1) v = vals[t.y >= 0 ? Bk.val_off + t.y : Bk.slot_base - t.y - 1];
    else {
      v = 0.0;
      for (int p = 0; p < cntv; ++p) {
        const int sidx = P.tgt_src[t.y + p];
        v += vals[sidx >= 0 ? Bk.val_off + sidx : Bk.slot_base - sidx - 1];
      }
    }
    int r = t.x & 255;
    if (r >= nc) r += nd_in;
    F[r + ((t.x >> 8) & 255) * ld] = v;
  }
  gsync<G>();
  PP_TRP(3);
  // children: fixed order (staged small ones in index order, then the others in index order)
  if (staged) {
    const int nbig = B.sh[5];
    for (int k = tid; k < H.nch; k += G) {
      const int e0 = stg.cnt[k], e1 = stg.cnt[k + 1];
      if (e1 == e0) continue;  // big child (or empty contribution)
      const int c = P.child_idx[H.ch0 + k];
      const SnHead C = P.heads[c];
      const int cne = Bk.meta[3 * c], ndo = Bk.meta[3 * c + 2];
      const int dim = Bk.meta[3 * c + 1] - cne;
      const int *ids = Bk.fid + C.fid_off + cne;
      const int *crel = P.rel + C.r0;
      const double *M = Bk.cb + C.cb_off;
      const int off = stg.ndo[k];
      // all global loads of a batch are issued before the first shared-memory store (the compiler cannot prove that
      // the staging area does not alias the contribution block, so interleaved loads would each pay a full round trip)
      int mp[SF_CHDIM];
#pragma unroll
      for (int i = 0; i < SF_CHDIM; ++i) mp[i] = (i < dim && i >= ndo) ? crel[i - ndo] : -1;
#pragma unroll
      for (int i = 0; i < SF_CHDIM; ++i) {
        if (i < dim) {
          if (i < ndo) { mp[i] = nc + off + i; B.fid[nc + off + i] = ids[i]; }
          else { const int rr = mp[i]; mp[i] = rr < nc ? rr : rr + nd_in; }
        }
      }
      int e = e0;
      if (dim <= 4) {
        double mv[10];
        int q = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = j; i < 4; ++i, ++q) mv[q] = i < dim ? M[i + j * dim] : 0.0;
        q = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = j; i < 4; ++i, ++q)
            if (i < dim) {
              const int a = mp[i] >= mp[j] ? mp[i] : mp[j], b = mp[i] >= mp[j] ? mp[j] : mp[i];
              stg.tgt[e] = a + b * ld;
              stg.val[e] = mv[q];
              ++e;
            }
      } else {
#pragma unroll
        for (int j = 0; j < SF_CHDIM; ++j) {
          if (j >= dim) break;
          double col[SF_CHDIM];
#pragma unroll
          for (int i = j; i < SF_CHDIM; ++i) col[i] = i < dim ? M[i + j * dim] : 0.0;
#pragma unroll
          for (int i = j; i < SF_CHDIM; ++i)
            if (i < dim) {
              const int a = mp[i] >= mp[j] ? mp[i] : mp[j], b = mp[i] >= mp[j] ? mp[j] : mp[i];
              stg.tgt[e] = a + b * ld;
              stg.val[e] = col[i];
              ++e;
            }
        }
      }
    }
    gsync<G>();
    PP_TRP(4);
    {
      const int total = stg.cnt[H.nch];
      apply_staged<G>(F, stg.tgt, stg.val, total);
    }
    gsync<G>();
    PP_TRP(5);
    for (int q = 0; q < nbig; ++q) {
      const int k = stg.big[q];
      const int c = P.child_idx[H.ch0 + k];
      const SnHead C = P.heads[c];
      const int cne = Bk.meta[3 * c], cS = Bk.meta[3 * c + 1], ndo = Bk.meta[3 * c + 2];
      const int dim = cS - cne, off = stg.ndo[k];
      const int *ids = Bk.fid + C.fid_off + cne;
      const int *crel = P.rel + C.r0;
      const double *M = Bk.cb + C.cb_off;
      for (int i = tid; i < dim; i += G) {
        if (i < ndo) { B.map[i] = nc + off + i; B.fid[nc + off + i] = ids[i]; }
        else { const int rr = crel[i - ndo]; B.map[i] = rr < nc ? rr : rr + nd_in; }
      }
      gsync<G>();
      if (dim <= LWG) {
        // one contribution row per lane: four columns of global loads in flight before the first update
        const int mi = lanep < dim ? B.map[lanep] : 0;
        for (int j0 = gwp; j0 < dim; j0 += 4 * NWG) {
          double mv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q * NWG;
            mv[q] = (j < dim && lanep >= j && lanep < dim) ? M[lanep + j * dim] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q * NWG;
            if (j < dim && lanep >= j && lanep < dim) fent(F, ld, mi, B.map[j]) += mv[q];
          }
        }
      } else {
        for (int j = gwp; j < dim; j += NWG) {
          const int mj = B.map[j];
          for (int i = j + lanep; i < dim; i += LWG) fent(F, ld, B.map[i], mj) += M[i + j * dim];
        }
      }
      gsync<G>();
    }
  } else {
    int off = 0;
    for (int k = 0; k < H.nch; ++k) {
      int cne, dim, ndo, c_fid, c_r0;
      long long c_cb;
      if (pref) {
        const unsigned gm = gmask<G>();
        constexpr int LWP = Grp<G>::LW;
        cne = __shfl_sync(gm, pcne, k, LWP);
        dim = __shfl_sync(gm, pdim, k, LWP);
        ndo = __shfl_sync(gm, pndo, k, LWP);
        c_fid = __shfl_sync(gm, pfid, k, LWP);
        c_r0 = __shfl_sync(gm, pr0, k, LWP);
        c_cb = __shfl_sync(gm, pcb, k, LWP);
      } else {
        const int c = P.child_idx[H.ch0 + k];
        const SnHead C = P.heads[c];
        cne = Bk.meta[3 * c];
        dim = Bk.meta[3 * c + 1] - cne;
        ndo = Bk.meta[3 * c + 2];
        c_fid = C.fid_off;
        c_r0 = C.r0;
        c_cb = C.cb_off;
      }
      const int *ids = Bk.fid + c_fid + cne;
      const int *crel = P.rel + c_r0;
      const double *M = Bk.cb + c_cb;
      for (int i = tid; i < dim; i += G) {
        if (i < ndo) { B.map[i] = nc + off + i; B.fid[nc + off + i] = ids[i]; }
        else { const int rr = crel[i - ndo]; B.map[i] = rr < nc ? rr : rr + nd_in; }
      }
      gsync<G>();
      if (dim <= LWG) {
        // one contribution row per lane: four columns of global loads in flight before the first update
        const int mi = lanep < dim ? B.map[lanep] : 0;
        for (int j0 = gwp; j0 < dim; j0 += 4 * NWG) {
          double mv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q * NWG;
            mv[q] = (j < dim && lanep >= j && lanep < dim) ? M[lanep + j * dim] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q * NWG;
            if (j < dim && lanep >= j && lanep < dim) fent(F, ld, mi, B.map[j]) += mv[q];
          }
        }
      } else {
        for (int j = gwp; j < dim; j += NWG) {
          const int mj = B.map[j];
          for (int i = j + lanep; i < dim; i += LWG) fent(F, ld, B.map[i], mj) += M[i + j * dim];
        }
      }
      gsync<G>();
      off += ndo;
    }
  }
  PP_TRP(6);
  int ne;
  if (G == SF_NT && S <= 48) {
    // small front assembled by the whole CTA: the pivot loop runs on four warps with a named barrier
    // (block-wide barriers of 16 warps would dominate it)
    if (tid < 128) {
      ne = factor_front<128>(B, S, fs, u, pivtol, cnt);
      if (tid == 0) B.sh[6] = ne;
    }
    __syncthreads();
    ne = B.sh[6];
  } else if (G == SF_MG && S <= 32) {
    // assembled by two warps, but a front this small pivots fastest on one (row-per-lane steps, warp barriers)
    if (tid < 32) {
      ne = factor_front<32>(B, S, fs, u, pivtol, cnt);
      if (tid == 0) B.sh[6] = ne;
    }
    gsync<G>();
    ne = B.sh[6];
  } else {
    ne = factor_front<G>(B, S, fs, u, pivtol, cnt);
  }
  PP_TRP(7);
  const int ndo = fs - ne, dim = S - ne;
  if (ndo > H.dslot) {
    if (tid == 0) {
      Bk.meta[3 * s] = 0; Bk.meta[3 * s + 1] = 0; Bk.meta[3 * s + 2] = 0;
      Bk.info[3] = 3; Bk.info[4] = s; Bk.info[5] = ndo;
    }
    return PF_FAIL;
  }
  const int caprows = nc + H.dcap + ncb;
  double *Ls = Bk.L + H.l_off;
  for (int j = gwp; j < ne; j += NWG)
    for (int i = j + lanep; i < S; i += LWG) Ls[i + (long long)j * caprows] = F[i + j * ld];
  for (int i = tid; i < S; i += G) Bk.fid[H.fid_off + i] = B.fid[i];
  for (int i = tid; i < fs; i += G) {
    Bk.opos[H.fs_off + i] = B.opos[i];
    if (i < ne) Bk.pbz[H.fs_off + i] = B.bsz[i];
  }
  double *M = Bk.cb + H.cb_off;
  for (int j = gwp; j < dim; j += NWG)
    for (int i = j + lanep; i < dim; i += LWG) M[i + j * dim] = F[(ne + i) + (ne + j) * ld];
  if (tid == 0) { Bk.meta[3 * s] = ne; Bk.meta[3 * s + 1] = S; Bk.meta[3 * s + 2] = ndo; }
  gsync<G>();
  PP_TRP(8);
  return PF_OK;
}

// Level-0 small fronts have no children, hence no dependencies at all: one warp per (block, leaf),
// spread over the whole GPU.  grid = (ceil(max leaves / LF_NW), blocks).
constexpr int LF_NT = 256, LF_NW = LF_NT / 32;

template <int LG>
__global__ void __launch_bounds__(LF_NT, 4) subtree_leaf_kernel(const SparseBlock *__restrict__ blocks,
                                                             const PlanDev *__restrict__ plans,
                                                             const double *__restrict__ vals, double u, double pivtol,
                                                             unsigned long long *inertia, int cap) {
  // cap = largest leaf front of any plan (leaves have no children, so their size is static).  LG lanes per leaf:
  // fronts of up to 8 (16) rows share a warp four (two) at a time, which divides the instruction count per leaf --
  // the leaf kernels are issue-bound -- and small per-leaf buffers keep 64 warps resident per SM.
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ int cnt[4];
  const SparseBlock Bk = blocks[blockIdx.y];
  const PlanDev P = plans[Bk.plan];
  constexpr int PER_CTA = LF_NT / LG;
  const int slot = threadIdx.x / LG;
  if (P.nlevels == 0) return;
  const int nleaf = P.tiny_ptr[1] - P.tiny_ptr[0];
  if ((int)blockIdx.x * PER_CTA >= nleaf) return;
  if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int k = blockIdx.x * PER_CTA + slot;
  if (k < nleaf) {
    const FrontBuf mine = carve(sm_raw + (size_t)slot * align16(fb_bytes(cap, cap | 1)), cap, cap | 1);
    Stage none;
    none.val = nullptr; none.tgt = nullptr; none.cnt = nullptr; none.ndo = nullptr; none.big = nullptr; none.cap = 0; none.maxch = 0;
    const int rc = process_front<LG>(Bk, P, vals, P.tiny_idx[P.tiny_ptr[0] + k], mine, u, pivtol, cnt, none);
    if (rc != PF_OK && gtid<LG>() == 0) Bk.info[2] = 1;
  }
  __syncthreads();
  if (threadIdx.x < 3 && cnt[threadIdx.x]) atomicAdd(&inertia[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

// One thread-block CLUSTER per block (1, 2, 4 or 8 CTAs, chosen by the host so that blocks x CTAs fills the GPU
// once): the fronts of a level are independent, so its tiny / medium / big fronts are dealt over the CTAs of the
// cluster; contribution blocks travel through HBM as before and a cluster barrier (release / acquire) closes each
// level.  With few blocks per GPU (strong scaling, or 64 blocks on 148 SMs) the levels with many fronts run on
// several SMs instead of one.
__global__ void __launch_bounds__(SF_NT) subtree_factor_kernel(const SparseBlock *__restrict__ blocks,
                                                               const PlanDev *__restrict__ plans,
                                                               Front *__restrict__ fronts,
                                                               const double *__restrict__ vals, double u,
                                                               double pivtol, unsigned long long *inertia) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), cr = (int)cl.block_rank();
  extern __shared__ __align__(16) unsigned char sm_raw[];
  int *cnt = reinterpret_cast<int *>(sm_raw + SF_WORK);  // [0..2] inertia, [3] failed, [4] deferred count
  int *deferred = cnt + 16;                              // up to 64 deferred fronts per level
  const SparseBlock Bk = blocks[blockIdx.x / C];
  const PlanDev P = plans[Bk.plan];
  const Front R = fronts[Bk.root];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = tid / SF_MG;
  // the three work areas alias each other: the passes of a level are separated by block barriers
  const FrontBuf big = carve(sm_raw, SF_SBUF, SF_LDF);
  const Stage stg = carve_stage(sm_raw + SF_BIG_FB, SF_STG, SF_MAXCH);
  const FrontBuf med = carve(sm_raw + (size_t)grp * SF_MED_BYTES, SF_MBUF, SF_MLD);
  const Stage mstg = carve_stage(sm_raw + (size_t)grp * SF_MED_BYTES + SF_MED_FB, SF_MSTG, SF_MMAXCH);
  const FrontBuf mine = carve(sm_raw + (size_t)warp * SF_TINY_BYTES, SF_TBUF, SF_TLD);
  if (tid < 8) cnt[tid] = 0;
  if (tid == 0 && cr == 0) { Bk.info[3] = 0; Bk.info[6] = 0; }
  __syncthreads();

  for (int l = 0; l < P.nlevels; ++l) {
    PP_TR(4 * l);
    // ---- small fronts: one warp each, all warps concurrently (the leaves were done by
    //      subtree_leaf_kernel, spread over the whole GPU) ----
    for (int k = P.tiny_ptr[l] + warp + SF_NW * cr; l > 0 && k < P.tiny_ptr[l + 1]; k += SF_NW * C) {
      const int s = P.tiny_idx[k];
      const int rc = process_front<32>(Bk, P, vals, s, mine, u, pivtol, cnt, mstg);
      if (lane == 0) {
        if (rc == PF_DEFER) {
          const int pos = atomicAdd(&cnt[4], 1);
          if (pos < 64) deferred[pos] = s; else cnt[3] = 1;
        } else if (rc == PF_FAIL) cnt[3] = 1;
      }
    }
    __syncthreads();
    PP_TR(4 * l + 1);
    // ---- medium fronts (up to 32 rows, any number of children): two warps each, eight at a time ----
    for (int k = P.med_ptr[l] + grp + SF_NG * cr; k < P.med_ptr[l + 1]; k += SF_NG * C) {
      const int s = P.med_idx[k];
      const int rc = process_front<SF_MG>(Bk, P, vals, s, med, u, pivtol, cnt, mstg);
      if ((tid & (SF_MG - 1)) == 0) {
        if (rc == PF_DEFER) {
          const int pos = atomicAdd(&cnt[4], 1);
          if (pos < 64) deferred[pos] = s; else cnt[3] = 1;
        } else if (rc == PF_FAIL) cnt[3] = 1;
      }
    }
    __syncthreads();
    PP_TR(4 * l + 2);
    // ---- larger fronts (and smaller ones that grew through delayed pivots): whole CTA, one by one ----
    const int ndef = min(cnt[4], 64);
    for (int k = 0; k < ndef; ++k) {
      int best = -1;  // deterministic order: ascending supernode index
      for (int q = 0; q < ndef; ++q)
        if (deferred[q] >= 0 && (best < 0 || deferred[q] < deferred[best])) best = q;
      const int s = deferred[best];
      __syncthreads();
      if (tid == 0) deferred[best] = -1;
      if (process_front<SF_NT>(Bk, P, vals, s, big, u, pivtol, cnt, stg) != PF_OK && tid == 0) cnt[3] = 1;
      __syncthreads();
    }
    for (int k = P.big_ptr[l] + cr; k < P.big_ptr[l + 1]; k += C) {
      if (process_front<SF_NT>(Bk, P, vals, P.big_idx[k], big, u, pivtol, cnt, stg) != PF_OK && tid == 0) cnt[3] = 1;
      __syncthreads();
    }
    if (tid == 0) cnt[4] = 0;
    // a failed front leaves its (structurally valid) previous record in place, so the other fronts can go on safely;
    // the failure is reported at the end -- no early exit that the CTAs of a cluster would have to agree on
    if (C > 1) cl.sync(); else __syncthreads();
    PP_TR(4 * l + 3);
  }
  if (tid == 0) {
    if (cnt[3]) atomicOr(&Bk.info[6], 1);
    if (cnt[0]) atomicAdd(&inertia[0], (unsigned long long)cnt[0]);
    if (cnt[1]) atomicAdd(&inertia[1], (unsigned long long)cnt[1]);
    if (cnt[2]) atomicAdd(&inertia[2], (unsigned long long)cnt[2]);
  }
  if (C > 1) { __threadfence(); cl.sync(); }
  if (cr != 0) return;
  // ---- children of the root: add their contribution blocks into the dense root front ----
  int ndroot = 0;
  bool failed = *((volatile int *)&Bk.info[6]) != 0 || cnt[3] != 0 || Bk.info[2] != 0;
  for (int k = 0; k < P.nrootch && !failed; ++k) {
    const int s = P.root_children[k];
    const int ne = Bk.meta[3 * s], S = Bk.meta[3 * s + 1], ndo = Bk.meta[3 * s + 2];
    const int dim = S - ne;
    if (ndroot + ndo > P.DR) {
      if (tid == 0) { Bk.info[3] = 4; Bk.info[4] = s; Bk.info[5] = ndroot + ndo; }
      failed = true;
      break;
    }
    const SnHead H = P.heads[s];
    const int *ids = Bk.fid + H.fid_off + ne;
    const int *crel = P.rel + H.r0;
    const double *M = Bk.cb + H.cb_off;
    for (int i = tid; i < dim; i += SF_NT) {
      if (i < ndo) { big.map[i] = P.nT + ndroot + i; Bk.rootids[P.nT + ndroot + i] = ids[i]; }
      else big.map[i] = crel[i - ndo];
    }
    __syncthreads();
    for (int idx = tid; idx < dim * dim; idx += SF_NT) {
      const int j = idx / dim, i = idx - j * dim;
      if (i < j) continue;
      int a = big.map[i], b = big.map[j];
      if (a < b) { const int x = a; a = b; b = x; }
      R.A[(size_t)a + (size_t)b * R.ld] += M[i + j * dim];
    }
    __syncthreads();
    ndroot += ndo;
  }
  for (int i = tid; i < P.nT; i += SF_NT) Bk.rootids[i] = P.rootcols[i];
  for (int t = ndroot + tid; t < P.DR; t += SF_NT) Bk.rootids[P.nT + t] = -1;  // unused delayed slots
  __syncthreads();
  if (tid == 0) {
    Bk.info[0] = failed ? 1 : 0;
    Bk.info[1] = ndroot;
    Bk.info[2] = 0;  // leaf-failure flag consumed
    PP_TR(4 * P.nlevels);
    fronts[Bk.root].n = P.nT + ndroot;  // pivot candidates of the dense root: static columns + delayed ones
  }
}

// ---- solves on the subtree part ---------------------------------------------------------------
// Per front (after pivoting): rows [0,ne) eliminated, [ne,fs) delayed, [fs,S) contribution rows.
//   forward : v = rhs on own rows + children's vectors;  z = L11^-1 v[0:ne];  out = v[ne:] - L21 z
//   backward: x[0:ne] = L11^-T (D^-1 z - L21^T x[ne:])
struct SolveBuf {
  double *Ls, *z, *v;
  int *fid, *inv, *bsz, *map;
  int ld, cap;
};
__host__ __device__ constexpr size_t sb_bytes(int cap, int ld) {
  return (size_t)cap * ld * 8 + 2 * (size_t)cap * 8 + 4 * (size_t)cap * 4;
}
constexpr size_t SV_TINY_BYTES = align16(sb_bytes(SF_TBUF, SF_TLD));
constexpr size_t SV_BIG_SB = align16(sb_bytes(SF_SBUF, SF_LDF));
constexpr size_t SV_MED_SB = align16(sb_bytes(SF_MBUF, SF_MLD));
constexpr size_t SV_MED_BYTES = SV_MED_SB + stage_bytes(SF_MSTG, SF_MMAXCH);
constexpr size_t SV_SMEM = max3(SV_BIG_SB + stage_bytes(SF_STG, SF_MAXCH), SF_NG * SV_MED_BYTES, SF_NW * SV_TINY_BYTES);

__device__ __forceinline__ SolveBuf carve_solve(unsigned char *base, int cap, int ld) {
  SolveBuf b;
  b.Ls = reinterpret_cast<double *>(base);
  b.z = b.Ls + (size_t)cap * ld;
  b.v = b.z + cap;
  b.fid = reinterpret_cast<int *>(b.v + cap);
  b.inv = b.fid + cap;
  b.bsz = b.inv + cap;
  b.map = b.bsz + cap;
  b.ld = ld;
  b.cap = cap;
  return b;
}

template <int G>
__device__ void load_front(const SparseBlock &Bk, const SnHead &H, const SolveBuf &B, int ne, int S) {
  const int tid = gtid<G>();
  const int caprows = H.nc + H.dcap + H.ncb;
  const double *Lg = Bk.L + H.l_off;
  constexpr int NWL = Grp<G>::NW, LWL = Grp<G>::LW;
  const int wl = G < 32 ? 0 : tid >> 5, ll = tid & (LWL - 1);
  if (S <= LWL) {
    // short columns: four columns of global loads in flight before the first shared-memory store
    for (int j0 = wl; j0 < ne; j0 += 4 * NWL) {
      double v4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + q * NWL;
        v4[q] = (j < ne && ll >= j && ll < S) ? Lg[ll + (long long)j * caprows] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + q * NWL;
        if (j < ne && ll >= j && ll < S) B.Ls[ll + j * B.ld] = v4[q];
      }
    }
  } else {
    for (int j = wl; j < ne; j += NWL)
      for (int i = j + ll; i < S; i += LWL) B.Ls[i + j * B.ld] = Lg[i + (long long)j * caprows];
  }
  for (int i = tid; i < S; i += G) B.fid[i] = Bk.fid[H.fid_off + i];
  for (int i = tid; i < ne; i += G) B.bsz[i] = Bk.pbz[H.fs_off + i];
}

template <int G>
__device__ void forward_front(const SparseBlock &Bk, const PlanDev &P, int s, const SolveBuf &B,
                              const double *__restrict__ rhs, double *__restrict__ y, const Stage &stg) {
  constexpr int LW = Grp<G>::LW;
  const int tid = gtid<G>(), lane = tid & (LW - 1);
  PP_TRP(0);
  const int ne = Bk.meta[3 * s], S = Bk.meta[3 * s + 1], fs = ne + Bk.meta[3 * s + 2];
  if (S == 0) return;
  const SnHead H = P.heads[s];
  const int nc = H.nc;
  const int nd_in = fs - nc;
  load_front<G>(Bk, H, B, ne, S);
  for (int i = tid; i < fs; i += G) B.inv[Bk.opos[H.fs_off + i]] = i;  // pre-pivot -> stored position
  gsync<G>();
  PP_TRP(1);
  for (int i = tid; i < S; i += G)
    B.v[i] = (i < fs && Bk.opos[H.fs_off + i] < nc) ? rhs[B.fid[i]] : 0.0;
  gsync<G>();
  PP_TRP(2);
  if (G >= SF_MG && H.nch >= 4 && H.nch <= stg.maxch) {
    // many children: fetch their vectors concurrently (one child per thread), apply in child order
    for (int k = tid; k < H.nch; k += G) {
      const int c = P.child_idx[H.ch0 + k];
      stg.ndo[k] = Bk.meta[3 * c + 2];
      stg.cnt[k] = Bk.meta[3 * c + 1] - Bk.meta[3 * c];
    }
    gsync<G>();
    if (tid < 32) stage_prefix_warp(stg, H.nch, stg.big);  // no big children here: big[] doubles as scratch
    gsync<G>();
    const int total = stg.cnt[H.nch];
    PP_TRP(3);
    if (total <= stg.cap) {
      for (int k = tid; k < H.nch; k += G) {
        const int c = P.child_idx[H.ch0 + k];
        const SnHead C = P.heads[c];
        const int ndo = Bk.meta[3 * c + 2], dim = stg.cnt[k + 1] - stg.cnt[k], off = stg.ndo[k];
        const int *crel = P.rel + C.r0;
        const double *uc = Bk.vec + C.vec_off;
        const int e0 = stg.cnt[k];
        for (int i0 = 0; i0 < dim; i0 += 8) {  // loads of a batch first, then the shared-memory stores
          int rl[8];
          double uv[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int i = i0 + r;
            rl[r] = (i < dim && i >= ndo) ? crel[i - ndo] : 0;
            uv[r] = i < dim ? uc[i] : 0.0;
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int i = i0 + r;
            if (i < dim) {
              int q;
              if (i < ndo) q = nc + off + i;
              else q = rl[r] < nc ? rl[r] : rl[r] + nd_in;
              stg.tgt[e0 + i] = q < fs ? B.inv[q] : q;
              stg.val[e0 + i] = uv[r];
            }
          }
        }
      }
      gsync<G>();
      PP_TRP(4);
      apply_staged<G>(B.v, stg.tgt, stg.val, total);
      gsync<G>();
      PP_TRP(5);
    } else {
      for (int k = 0; k < H.nch; ++k) {
        const int c = P.child_idx[H.ch0 + k];
        const SnHead C = P.heads[c];
        const int ndo = Bk.meta[3 * c + 2], dim = stg.cnt[k + 1] - stg.cnt[k], off = stg.ndo[k];
        const int *crel = P.rel + C.r0;
        const double *uc = Bk.vec + C.vec_off;
        for (int i = tid; i < dim; i += G) {
          int q;
          if (i < ndo) q = nc + off + i;
          else { const int rr = crel[i - ndo]; q = rr < nc ? rr : rr + nd_in; }
          B.v[q < fs ? B.inv[q] : q] += uc[i];
        }
        gsync<G>();
      }
    }
  } else {
    int off = 0;
    for (int k = 0; k < H.nch; ++k) {
      const int c = P.child_idx[H.ch0 + k];
      const SnHead C = P.heads[c];
      const int cne = Bk.meta[3 * c], cS = Bk.meta[3 * c + 1], ndo = Bk.meta[3 * c + 2];
      const int dim = cS - cne;
      const int *crel = P.rel + C.r0;
      const double *uc = Bk.vec + C.vec_off;
      for (int i = tid; i < dim; i += G) {
        int q;
        if (i < ndo) q = nc + off + i;
        else { const int rr = crel[i - ndo]; q = rr < nc ? rr : rr + nd_in; }
        B.v[q < fs ? B.inv[q] : q] += uc[i];
      }
      gsync<G>();
      off += ndo;
    }
  }
  if (tid < 32) {
    if (ne <= LW) {
      // one row per lane: the solved entries travel by shuffle, the L entries are independent loads that pipeline
      // (the loop over shared memory below costs a load-FMA-store-barrier chain per column)
      const unsigned wm = gmask<G>();
      double x = lane < ne ? B.v[lane] : 0.0;
      const int flag = lane < ne ? B.bsz[lane] : 1;
      for (int c = 0; c < ne; ++c) {
        const double zc = __shfl_sync(wm, x, c, LW);
        const int fc = __shfl_sync(wm, flag, c, LW);
        if (lane > c && lane < ne && !(fc == 2 && lane == c + 1)) x -= B.Ls[lane + c * B.ld] * zc;
      }
      if (lane < ne) B.v[lane] = x;
    } else {
      for (int c = 0; c < ne; ++c) {
        const double zc = B.v[c];
        const int skip = B.bsz[c] == 2 ? c + 1 : -1;
        for (int i = c + 1 + lane; i < ne; i += LW)
          if (i != skip) B.v[i] -= B.Ls[i + c * B.ld] * zc;
        __syncwarp(gmask<G>());
      }
    }
  }
  gsync<G>();
  PP_TRP(6);
  double *out = Bk.vec + H.vec_off;
  for (int i = ne + tid; i < S; i += G) {
    double acc = 0.0;
    for (int c = 0; c < ne; ++c) acc += B.Ls[i + c * B.ld] * B.v[c];
    out[i - ne] = B.v[i] - acc;
  }
  for (int i = tid; i < ne; i += G) y[B.fid[i]] = B.v[i];
  gsync<G>();
  PP_TRP(7);
}

template <int G>
__device__ void backward_front(const SparseBlock &Bk, const PlanDev &P, int s, const SolveBuf &B,
                               const double *__restrict__ y, double *__restrict__ x) {
  constexpr int LW = Grp<G>::LW;
  const int tid = gtid<G>(), lane = tid & (LW - 1);
  const int ne = Bk.meta[3 * s], S = Bk.meta[3 * s + 1];
  if (ne == 0) return;
  const SnHead H = P.heads[s];
  load_front<G>(Bk, H, B, ne, S);
  gsync<G>();
  for (int i = ne + tid; i < S; i += G) B.v[i] = x[B.fid[i]];
  for (int k = tid; k < ne; k += G) {  // w = D^-1 z
    const int b = B.bsz[k];
    if (b == 1) {
      const double d = B.Ls[k + k * B.ld];
      B.z[k] = d != 0.0 ? y[B.fid[k]] / d : 0.0;
    } else if (b == 2) {
      const double e21 = B.Ls[k + 1 + k * B.ld];
      const double akm1 = B.Ls[k + k * B.ld] / e21, ak = B.Ls[k + 1 + (k + 1) * B.ld] / e21;
      const double denom = akm1 * ak - 1.0;
      const double bkm1 = y[B.fid[k]] / e21, bk = y[B.fid[k + 1]] / e21;
      B.z[k] = (ak * bkm1 - bk) / denom;
      B.z[k + 1] = (akm1 * bk - bkm1) / denom;
    }
  }
  gsync<G>();
  for (int c = tid; c < ne; c += G) {
    double acc = 0.0;
    for (int i = ne; i < S; ++i) acc += B.Ls[i + c * B.ld] * B.v[i];
    B.z[c] -= acc;
  }
  gsync<G>();
  if (tid < 32) {
    if (ne <= LW) {
      const unsigned wm = gmask<G>();
      double x = lane < ne ? B.z[lane] : 0.0;
      const bool pair = lane < ne && B.bsz[lane] == 2;  // (lane, lane + 1) is a 2x2 pivot: no L entry between them
      for (int rr = ne - 1; rr > 0; --rr) {
        const double xv = __shfl_sync(wm, x, rr, LW);
        if (lane < rr && !(pair && rr == lane + 1)) x -= B.Ls[rr + lane * B.ld] * xv;
      }
      if (lane < ne) B.z[lane] = x;
    } else {
      for (int rr = ne - 1; rr > 0; --rr) {
        const double xv = B.z[rr];
        for (int c = lane; c < rr; c += LW)
          if (!(B.bsz[c] == 2 && rr == c + 1)) B.z[c] -= B.Ls[rr + c * B.ld] * xv;
        __syncwarp(gmask<G>());
      }
    }
  }
  gsync<G>();
  for (int i = tid; i < ne; i += G) x[B.fid[i]] = B.z[i];
  gsync<G>();
}

__global__ void __launch_bounds__(SF_NT) subtree_forward_kernel(const SparseBlock *__restrict__ blocks,
                                                                const PlanDev *__restrict__ plans,
                                                                const double *__restrict__ rhs,
                                                                const long long *__restrict__ vec_off,
                                                                double *__restrict__ ywork,
                                                                double *__restrict__ root_rhs,
                                                                const long long *__restrict__ root_off) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), cr = (int)cl.block_rank(), bid = blockIdx.x / C;  // one cluster per block
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const SparseBlock Bk = blocks[bid];
  const PlanDev P = plans[Bk.plan];
  const int tid = threadIdx.x, warp = tid >> 5, grp = tid / SF_MG;
  const SolveBuf big = carve_solve(sm_raw, SF_SBUF, SF_LDF);
  const Stage stg = carve_stage(sm_raw + SV_BIG_SB, SF_STG, SF_MAXCH);
  const SolveBuf med = carve_solve(sm_raw + (size_t)grp * SV_MED_BYTES, SF_MBUF, SF_MLD);
  const Stage mstg = carve_stage(sm_raw + (size_t)grp * SV_MED_BYTES + SV_MED_SB, SF_MSTG, SF_MMAXCH);
  const SolveBuf mine = carve_solve(sm_raw + (size_t)warp * SV_TINY_BYTES, SF_TBUF, SF_TLD);
  const double *r = rhs + vec_off[bid];
  double *y = ywork + vec_off[bid];
  for (int l = 0; l < P.nlevels; ++l) {
    PP_TR(1024 + 4 * l);
    for (int k = P.tiny_ptr[l] + warp + SF_NW * cr; l > 0 && k < P.tiny_ptr[l + 1]; k += SF_NW * C) {
      const int s = P.tiny_idx[k];
      if (Bk.meta[3 * s + 1] <= SF_TBUF) forward_front<32>(Bk, P, s, mine, r, y, mstg);
    }
    __syncthreads();
    PP_TR(1024 + 4 * l + 1);
    for (int k = P.med_ptr[l] + grp + SF_NG * cr; k < P.med_ptr[l + 1]; k += SF_NG * C) {
      const int s = P.med_idx[k];
      if (Bk.meta[3 * s + 1] <= SF_MBUF) forward_front<SF_MG>(Bk, P, s, med, r, y, mstg);
    }
    __syncthreads();
    PP_TR(1024 + 4 * l + 2);
    // whole-CTA fronts of the level (those that outgrew their group, and the big ones), dealt over the cluster
    int turn = 0;
    for (int k = P.tiny_ptr[l]; l > 0 && k < P.tiny_ptr[l + 1]; ++k) {  // fronts that outgrew their group
      const int s = P.tiny_idx[k];
      if (Bk.meta[3 * s + 1] > SF_TBUF && (turn++ % C) == cr) { forward_front<SF_NT>(Bk, P, s, big, r, y, stg); __syncthreads(); }
    }
    for (int k = P.med_ptr[l]; k < P.med_ptr[l + 1]; ++k) {
      const int s = P.med_idx[k];
      if (Bk.meta[3 * s + 1] > SF_MBUF && (turn++ % C) == cr) { forward_front<SF_NT>(Bk, P, s, big, r, y, stg); __syncthreads(); }
    }
    for (int k = P.big_ptr[l]; k < P.big_ptr[l + 1]; ++k) {
      if ((turn++ % C) != cr) continue;
      forward_front<SF_NT>(Bk, P, P.big_idx[k], big, r, y, stg);
      __syncthreads();
    }
    if (C > 1) cl.sync(); else __syncthreads();
    PP_TR(1024 + 4 * l + 3);
  }
  if (cr != 0) return;
  // root right-hand side: own entries + contributions of the root's children (fixed order)
  double *rr = root_rhs + root_off[bid];
  for (int p = tid; p < P.nT + P.DR; p += SF_NT) rr[p] = p < P.nT ? r[P.rootcols[p]] : 0.0;
  __syncthreads();
  int ndroot = 0;
  for (int k = 0; k < P.nrootch; ++k) {
    const int s = P.root_children[k];
    const int ne = Bk.meta[3 * s], S = Bk.meta[3 * s + 1], ndo = Bk.meta[3 * s + 2];
    const int dim = S - ne;
    const SnHead H = P.heads[s];
    const int *crel = P.rel + H.r0;
    const double *uc = Bk.vec + H.vec_off;
    for (int i = tid; i < dim; i += SF_NT) rr[i < ndo ? P.nT + ndroot + i : crel[i - ndo]] += uc[i];
    __syncthreads();
    ndroot += ndo;
  }
}

__global__ void __launch_bounds__(SF_NT) subtree_backward_kernel(const SparseBlock *__restrict__ blocks,
                                                                 const PlanDev *__restrict__ plans,
                                                                 const double *__restrict__ ywork,
                                                                 const long long *__restrict__ vec_off,
                                                                 const double *__restrict__ root_x,
                                                                 const long long *__restrict__ root_off,
                                                                 double *__restrict__ xout) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks(), cr = (int)cl.block_rank(), bid = blockIdx.x / C;  // one cluster per block
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const SparseBlock Bk = blocks[bid];
  const PlanDev P = plans[Bk.plan];
  const int tid = threadIdx.x, warp = tid >> 5, grp = tid / SF_MG;
  const SolveBuf big = carve_solve(sm_raw, SF_SBUF, SF_LDF);
  const SolveBuf med = carve_solve(sm_raw + (size_t)grp * SV_MED_BYTES, SF_MBUF, SF_MLD);
  const SolveBuf mine = carve_solve(sm_raw + (size_t)warp * SV_TINY_BYTES, SF_TBUF, SF_TLD);
  const double *y = ywork + vec_off[bid];
  double *x = xout + vec_off[bid];
  const double *rx = root_x + root_off[bid];
  if (cr == 0)
    for (int p = tid; p < P.nT + P.DR; p += SF_NT) {
      const int id = Bk.rootids[p];
      if (id >= 0) x[id] = rx[p];
    }
  if (C > 1) { __threadfence(); cl.sync(); } else __syncthreads();
  for (int l = P.nlevels - 1; l >= 0; --l) {
    int turn = 0;  // whole-CTA fronts of the level are dealt over the CTAs of the cluster
    for (int k = P.big_ptr[l]; k < P.big_ptr[l + 1]; ++k) {
      if ((turn++ % C) != cr) continue;
      backward_front<SF_NT>(Bk, P, P.big_idx[k], big, y, x);
      __syncthreads();
    }
    for (int k = P.med_ptr[l]; k < P.med_ptr[l + 1]; ++k) {
      const int s = P.med_idx[k];
      if (Bk.meta[3 * s + 1] > SF_MBUF && (turn++ % C) == cr) { backward_front<SF_NT>(Bk, P, s, big, y, x); __syncthreads(); }
    }
    for (int k = P.tiny_ptr[l]; l > 0 && k < P.tiny_ptr[l + 1]; ++k) {
      const int s = P.tiny_idx[k];
      if (Bk.meta[3 * s + 1] > SF_TBUF && (turn++ % C) == cr) { backward_front<SF_NT>(Bk, P, s, big, y, x); __syncthreads(); }
    }
    __syncthreads();
    for (int k = P.med_ptr[l] + grp + SF_NG * cr; k < P.med_ptr[l + 1]; k += SF_NG * C) {
      const int s = P.med_idx[k];
      if (Bk.meta[3 * s + 1] <= SF_MBUF) backward_front<SF_MG>(Bk, P, s, med, y, x);
    }
    __syncthreads();
    for (int k = P.tiny_ptr[l] + warp + SF_NW * cr; l > 0 && k < P.tiny_ptr[l + 1]; k += SF_NW * C) {
      const int s = P.tiny_idx[k];
      if (Bk.meta[3 * s + 1] <= SF_TBUF) backward_front<32>(Bk, P, s, mine, y, x);
    }
    if (C > 1) cl.sync(); else __syncthreads();
  }
}

// leaves of the solves: one warp per (block, leaf) over the whole GPU

template <int LG>
__global__ void __launch_bounds__(LF_NT) subtree_leaf_forward_kernel(const SparseBlock *__restrict__ blocks,
                                                                     const PlanDev *__restrict__ plans,
                                                                     const double *__restrict__ rhs,
                                                                     const long long *__restrict__ vec_off,
                                                                     double *__restrict__ ywork, int cap) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const SparseBlock Bk = blocks[blockIdx.y];
  const PlanDev P = plans[Bk.plan];
  const int slot = threadIdx.x / LG;
  if (P.nlevels == 0) return;
  const int k = blockIdx.x * (LF_NT / LG) + slot;
  if (k >= P.tiny_ptr[1] - P.tiny_ptr[0]) return;
  const SolveBuf mine = carve_solve(sm_raw + (size_t)slot * align16(sb_bytes(cap, cap | 1)), cap, cap | 1);
  Stage none;
  none.val = nullptr; none.tgt = nullptr; none.cnt = nullptr; none.ndo = nullptr; none.big = nullptr; none.cap = 0; none.maxch = 0;
  forward_front<LG>(Bk, P, P.tiny_idx[P.tiny_ptr[0] + k], mine, rhs + vec_off[blockIdx.y], ywork + vec_off[blockIdx.y], none);
}

template <int LG>
__global__ void __launch_bounds__(LF_NT) subtree_leaf_backward_kernel(const SparseBlock *__restrict__ blocks,
                                                                      const PlanDev *__restrict__ plans,
                                                                      const double *__restrict__ ywork,
                                                                      const long long *__restrict__ vec_off,
                                                                      double *__restrict__ xout, int cap) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const SparseBlock Bk = blocks[blockIdx.y];
  const PlanDev P = plans[Bk.plan];
  const int slot = threadIdx.x / LG;
  if (P.nlevels == 0) return;
  const int k = blockIdx.x * (LF_NT / LG) + slot;
  if (k >= P.tiny_ptr[1] - P.tiny_ptr[0]) return;
  const SolveBuf mine = carve_solve(sm_raw + (size_t)slot * align16(sb_bytes(cap, cap | 1)), cap, cap | 1);
  backward_front<LG>(Bk, P, P.tiny_idx[P.tiny_ptr[0] + k], mine, ywork + vec_off[blockIdx.y], xout + vec_off[blockIdx.y]);
}

// End of the local phase in ONE launch (one CTA per block): the inertia of every dense root front is added to the
// counters, and the CTA that finishes last (ticket in inertia[7]; no waiting) reduces the status flags of all fronts
// and sparse blocks -- flag[0] = a root reported a zero pivot, flag[1] = a sparse block ran out of delayed-pivot
// capacity -- and packs the tail of the Schur buffer (status + inertia) that the caller all-reduces.
__global__ void local_finalize_kernel(const Front *__restrict__ fronts, const SparseBlock *__restrict__ blocks,
                                      int count, unsigned long long *inertia, int *flag, double *tail) {
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
      const double det = (a / o) * (c / o) - 1.0;  // sign of a*c - o^2, scaled as in the factorisation
      if (det < 0.0) { pos += 1; neg += 1; }
      else if (det > 0.0) { if (a + c > 0.0) pos += 2; else neg += 2; }
      else { zero += 1; if (a + c > 0.0) pos += 1; else if (a + c < 0.0) neg += 1; else zero += 1; }
    }
  }
  __shared__ int s[3];
  __shared__ int last;
  if (threadIdx.x < 3) s[threadIdx.x] = 0;
  __syncthreads();
  atomicAdd(&s[0], pos);
  atomicAdd(&s[1], neg);
  atomicAdd(&s[2], zero);
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(&inertia[threadIdx.x], (unsigned long long)s[threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&inertia[7], 1ull) == (unsigned long long)(gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  int bad = 0, sbad = 0;
  for (int f = threadIdx.x; f < count; f += blockDim.x) {
    if (fronts[f].state[ST_INFO] != 0) bad = 1;
    if (blocks[f].info[0] != 0) sbad = 1;
  }
  bad = __syncthreads_or(bad);
  sbad = __syncthreads_or(sbad);
  if (threadIdx.x == 0) {
    flag[0] = bad;
    flag[1] = sbad;
    if (tail) {
      tail[0] = bad ? 1.0 : 0.0;
      tail[1] = sbad ? 1.0 : 0.0;
      for (int k = 0; k < 3; ++k) tail[2 + k] = (double)*((volatile unsigned long long *)&inertia[k]);
      tail[5] = tail[6] = tail[7] = 0.0;
    }
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
