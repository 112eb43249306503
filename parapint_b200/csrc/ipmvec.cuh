// Interior-point vector kernels: the O(n) passes parapint's IPM loop makes over its iterates
// between two linear solves, each fused into ONE streaming pass with the reductions finished on the device.
//
//   fraction_to_the_boundary      algorithms/interior_point.py:655-758 (helpers :655-674), with the bound-multiplier
//                                 steps of interfaces/interface.py:548-570 formed on the fly instead of materialised
//   complementarity / scaling     algorithms/interior_point.py:241-251 (bound residuals), :274-315 (maxima, |dual|
//                                 sums, finite-bound counts)
//   max |a - b|                   :253-269 (primal / dual infeasibility maxima)
//   step update                   :587-626 (x += alpha * dx per iterate vector)
//
// All of it is HBM-bound: every vector is read exactly once with 16-byte loads, grid = a multiple of the SM count,
// grid-stride loop.  Arithmetic follows the reference's NumPy expressions operation by operation (no fused
// multiply-add: __dmul_rn / __dadd_rn / __dsub_rn / __ddiv_rn), so minima and maxima are BIT-IDENTICAL to the
// reference (they do not depend on the order of the reduction); sums are accumulated in a fixed order (thread-strided
// partials, warp tree, CTA tree, CTAs in index order by the last CTA to finish) and are reproducible run to run.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace ppb {

constexpr int IV_NT = 256;          // threads per CTA
constexpr int IV_MAXCTA = 148 * 8;  // partials per kernel (workspace rows)
constexpr int IV_SLOTS = 8;         // values reduced per kernel (workspace columns)

enum { IV_MIN = 0, IV_MAX = 1, IV_SUM = 2 };

__device__ __forceinline__ double iv_combine(int op, double a, double b) {
  return op == IV_MIN ? fmin(a, b) : op == IV_MAX ? fmax(a, b) : __dadd_rn(a, b);
}

// CTA-wide reduction of `NV` values per thread (ops[v] each), then the cross-CTA step: partials to `ws`
// ([gridDim.x][IV_SLOTS]), and the last CTA to arrive (ticket in `counter`) folds them in CTA order into out[v]
// with `accumulate[v]` deciding whether out[v] is combined with what it held (min / max over several groups of
// vectors) or overwritten.
template <int NV>
__device__ __forceinline__ void iv_finish(double (&val)[NV], const int (&ops)[NV], double *__restrict__ ws,
                                          unsigned int *__restrict__ counter, double *__restrict__ out, int accumulate) {
  __shared__ double sred[IV_NT / 32][NV];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) val[v] = iv_combine(ops[v], val[v], __shfl_down_sync(0xffffffffu, val[v], o));
    if (lane == 0) sred[warp][v] = val[v];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double a = sred[0][v];
      for (int w = 1; w < IV_NT / 32; ++w) a = iv_combine(ops[v], a, sred[w][v]);
      ws[(size_t)blockIdx.x * IV_SLOTS + v] = a;
    }
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last CTA folds the partials: thread t takes CTAs t, t + 256, ... in that order, then the same warp / CTA tree
  // as above -- a fixed association of the sums whatever order the CTAs finished in
  const volatile double *w = ws;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double a = ops[v] == IV_MIN ? CUDART_INF : ops[v] == IV_MAX ? -CUDART_INF : 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += IV_NT) a = iv_combine(ops[v], a, w[(size_t)b * IV_SLOTS + v]);
#pragma unroll
    for (int o = 16; o; o >>= 1) a = iv_combine(ops[v], a, __shfl_down_sync(0xffffffffu, a, o));
    if (lane == 0) sred[warp][v] = a;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    const int v = threadIdx.x;
    double a = sred[0][v];
    for (int q = 1; q < IV_NT / 32; ++q) a = iv_combine(ops[v], a, sred[q][v]);
    out[v] = accumulate ? iv_combine(ops[v], out[v], a) : a;
  }
  if (threadIdx.x == 0) *counter = 0;  // ready for the next kernel on the stream
}

// ---------------------------------------------------------------------------------------------
// fraction to the boundary of one group of variables (primals or slacks) with their bound multipliers:
//   out[0] = min(out[0], alpha over x  against lb and ub)                      (:655-674 with x, dx)
//   out[1] = min(out[1], alpha over zl, zu against 0 with the steps of interface.py:548-570)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double iv_ftb_lb(double tau, double x, double dx, double xl) {
  // alpha = -tau * (x - xl) / dx_mod;  alpha[dx >= 0] = inf      (dx_mod = dx where dx != 0)
  if (!(dx < 0.0)) return CUDART_INF;      // (a NaN step yields inf here, NaN in NumPy: not a state the loop survives)
  return __ddiv_rn(__dmul_rn(-tau, __dsub_rn(x, xl)), dx);
}
__device__ __forceinline__ double iv_ftb_ub(double tau, double x, double dx, double xu) {
  if (!(dx > 0.0)) return CUDART_INF;
  return __ddiv_rn(__dmul_rn(tau, __dsub_rn(xu, x)), dx);
}

__global__ void __launch_bounds__(IV_NT) ipm_ftb_kernel(int64_t n, double tau, double barrier,
                                                        const double *__restrict__ x, const double *__restrict__ dx,
                                                        const double *__restrict__ lb, const double *__restrict__ ub,
                                                        const double *__restrict__ zl, const double *__restrict__ zu,
                                                        double *__restrict__ ws, unsigned int *__restrict__ counter,
                                                        double *__restrict__ out, int vec) {
  double v[2] = {CUDART_INF, CUDART_INF};
  const int64_t stride = (int64_t)gridDim.x * IV_NT;
  auto one = [&](double xi, double di, double li, double ui, double zli, double zui) {
    v[0] = fmin(v[0], fmin(iv_ftb_lb(tau, xi, di, li), iv_ftb_ub(tau, xi, di, ui)));
    // multiplier steps, interface.py:548-570:  ((barrier -+ z * dx) / (x - lb | ub - x)) - z
    const double dzl = __dsub_rn(__ddiv_rn(__dsub_rn(barrier, __dmul_rn(zli, di)), __dsub_rn(xi, li)), zli);
    const double dzu = __dsub_rn(__ddiv_rn(__dadd_rn(barrier, __dmul_rn(zui, di)), __dsub_rn(ui, xi)), zui);
    v[1] = fmin(v[1], fmin(iv_ftb_lb(tau, zli, dzl, 0.0), iv_ftb_lb(tau, zui, dzu, 0.0)));
  };
  const int64_t n2 = vec ? n >> 1 : 0;   // 16-byte loads when every pointer is 16-byte aligned
  for (int64_t i = (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n2; i += stride) {
    const double2 X = reinterpret_cast<const double2 *>(x)[i], D = reinterpret_cast<const double2 *>(dx)[i];
    const double2 L = reinterpret_cast<const double2 *>(lb)[i], U = reinterpret_cast<const double2 *>(ub)[i];
    const double2 ZL = reinterpret_cast<const double2 *>(zl)[i], ZU = reinterpret_cast<const double2 *>(zu)[i];
    one(X.x, D.x, L.x, U.x, ZL.x, ZU.x);
    one(X.y, D.y, L.y, U.y, ZL.y, ZU.y);
  }
  for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n; i += stride) one(x[i], dx[i], lb[i], ub[i], zl[i], zu[i]);
  const int ops[2] = {IV_MIN, IV_MIN};
  iv_finish<2>(v, ops, ws, counter, out, 1);
}

// ---------------------------------------------------------------------------------------------
// complementarity and scaling terms of one group (:241-251, :274-315):
//   out[0] = max(out[0], max |(x - lb) zl - barrier| over finite lb)     out[1] likewise for ub
//   out[2] += sum |zl|     out[3] += sum |zu|     out[4] += #finite lb     out[5] += #finite ub
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IV_NT) ipm_compl_kernel(int64_t n, double barrier, const double *__restrict__ x,
                                                          const double *__restrict__ lb, const double *__restrict__ ub,
                                                          const double *__restrict__ zl, const double *__restrict__ zu,
                                                          double *__restrict__ ws, unsigned int *__restrict__ counter,
                                                          double *__restrict__ out, int vec) {
  double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * IV_NT;
  auto one = [&](double xi, double li, double ui, double zli, double zui) {
    if (li != -CUDART_INF) {    // np.isneginf(lb) entries are zeroed (:246) -- +inf or NaN lower bounds are not
      v[0] = fmax(v[0], fabs(__dsub_rn(__dmul_rn(__dsub_rn(xi, li), zli), barrier)));
    }
    if (!isinf(ui)) v[1] = fmax(v[1], fabs(__dsub_rn(__dmul_rn(__dsub_rn(ui, xi), zui), barrier)));
    v[2] = __dadd_rn(v[2], fabs(zli));
    v[3] = __dadd_rn(v[3], fabs(zui));
    v[4] += isfinite(li) ? 1.0 : 0.0;
    v[5] += isfinite(ui) ? 1.0 : 0.0;
  };
  const int64_t n2 = vec ? n >> 1 : 0;
  for (int64_t i = (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n2; i += stride) {
    const double2 X = reinterpret_cast<const double2 *>(x)[i];
    const double2 L = reinterpret_cast<const double2 *>(lb)[i], U = reinterpret_cast<const double2 *>(ub)[i];
    const double2 ZL = reinterpret_cast<const double2 *>(zl)[i], ZU = reinterpret_cast<const double2 *>(zu)[i];
    one(X.x, L.x, U.x, ZL.x, ZU.x);
    one(X.y, L.y, U.y, ZL.y, ZU.y);
  }
  for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n; i += stride) one(x[i], lb[i], ub[i], zl[i], zu[i]);
  const int ops[6] = {IV_MAX, IV_MAX, IV_SUM, IV_SUM, IV_SUM, IV_SUM};
  iv_finish<6>(v, ops, ws, counter, out, 1);
}

// out[0] = max(out[0], max |a - b|)  (b may be null: max |a|);  out[1] += sum |a|      (:253-269, :296-300)
__global__ void __launch_bounds__(IV_NT) ipm_maxabs_kernel(int64_t n, const double *__restrict__ a,
                                                           const double *__restrict__ b, double *__restrict__ ws,
                                                           unsigned int *__restrict__ counter, double *__restrict__ out,
                                                           int vec) {
  double v[2] = {0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * IV_NT;
  const int64_t n2 = vec ? n >> 1 : 0;
  for (int64_t i = (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n2; i += stride) {
    const double2 A = reinterpret_cast<const double2 *>(a)[i];
    double2 B = make_double2(0.0, 0.0);
    if (b) B = reinterpret_cast<const double2 *>(b)[i];
    v[0] = fmax(v[0], fmax(fabs(b ? __dsub_rn(A.x, B.x) : A.x), fabs(b ? __dsub_rn(A.y, B.y) : A.y)));
    v[1] = __dadd_rn(v[1], __dadd_rn(fabs(A.x), fabs(A.y)));
  }
  for (int64_t i = 2 * n2 + (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n; i += stride) {
    v[0] = fmax(v[0], fabs(b ? __dsub_rn(a[i], b[i]) : a[i]));
    v[1] = __dadd_rn(v[1], fabs(a[i]));
  }
  const int ops[2] = {IV_MAX, IV_SUM};
  iv_finish<2>(v, ops, ws, counter, out, 1);
}

// Step of one group (:587-626 with interface.py:548-570): the multiplier steps are formed from the values BEFORE the
// update, then  x += ls * (alpha_p dx),  zl += ls * (alpha_d dzl),  zu += ls * (alpha_d dzu)  -- the reference scales the
// steps by the fraction-to-the-boundary lengths first (:588-595) and by the line-search step second (:619-626).
// alpha = [alpha_primal, alpha_dual, ls] is read from DEVICE memory (what ipm_ftb_kernel left there).
__global__ void __launch_bounds__(IV_NT) ipm_step_kernel(int64_t n, const double *__restrict__ alpha, double barrier,
                                                         double *__restrict__ x, const double *__restrict__ dx,
                                                         const double *__restrict__ lb, const double *__restrict__ ub,
                                                         double *__restrict__ zl, double *__restrict__ zu) {
  const double ap = alpha[0], ad = alpha[1], ls = alpha[2];
  const int64_t stride = (int64_t)gridDim.x * IV_NT;
  for (int64_t i = (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n; i += stride) {
    const double xi = x[i], di = dx[i], li = lb[i], ui = ub[i], zli = zl[i], zui = zu[i];
    const double dzl = __dsub_rn(__ddiv_rn(__dsub_rn(barrier, __dmul_rn(zli, di)), __dsub_rn(xi, li)), zli);
    const double dzu = __dsub_rn(__ddiv_rn(__dadd_rn(barrier, __dmul_rn(zui, di)), __dsub_rn(ui, xi)), zui);
    x[i] = __dadd_rn(xi, __dmul_rn(ls, __dmul_rn(ap, di)));
    zl[i] = __dadd_rn(zli, __dmul_rn(ls, __dmul_rn(ad, dzl)));
    zu[i] = __dadd_rn(zui, __dmul_rn(ls, __dmul_rn(ad, dzu)));
  }
}

// y += ls * (alpha[which] * dy)   (equality / inequality multipliers, :590-591, :621-622)
__global__ void __launch_bounds__(IV_NT) ipm_axpy_kernel(int64_t n, const double *__restrict__ alpha, int which,
                                                         double *__restrict__ y, const double *__restrict__ dy) {
  const double a = alpha[which], ls = alpha[2];
  const int64_t stride = (int64_t)gridDim.x * IV_NT;
  for (int64_t i = (int64_t)blockIdx.x * IV_NT + threadIdx.x; i < n; i += stride)
    y[i] = __dadd_rn(y[i], __dmul_rn(ls, __dmul_rn(a, dy[i])));
}

__global__ void ipm_fill_kernel(double *__restrict__ out, int n, double value) {
  if ((int)threadIdx.x < n) out[threadIdx.x] = value;
}

}  // namespace ppb
