// Front descriptors shared by every kernel of the library.
//
// A "front" is one diagonal KKT block K_i augmented with the nonzero rows of its border A_i:
//
//        [ K_i   .  ]   n  rows   (pivots are chosen here only)
//    F = [ A_i   0  ]   m  rows   (ride along; become L_A = A P^T L^-T D^-1)
//
// stored column-major, lower triangle, leading dimension ld (multiple of 16 doubles).  After the
// n pivots are eliminated the trailing m x m block holds  -A_i K_i^-1 A_i^T, this block's whole
// contribution to the Schur complement (reference explicit_schur_complement.py:114-121 forms the
// same quantity with one leaf back-solve per border row).  The coupling matrix S is a front with
// m = 0.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ppb {

struct Front {
  double *A;     // nf x nf, column-major, lower triangle significant
  double *W;     // nf x NBMAX panel workspace (columns of L*D of the current panel), same ld
  double *zbuf;  // n: D^-1 L^-1 P r kept between the forward and backward solve phases
  double *bvec;  // m: border part of the forward solve (-L_A z)
  int *ipiv;     // n: row interchanged with the pivot row (fully-permuted convention)
  int *bsz;      // n: 1 = 1x1 pivot, 2 = first column of a 2x2 pivot, 0 = second column
  int *perm;     // n: perm[i] = original row now at position i   (P K P^T = L D L^T)
  int *state;    // [0] columns eliminated, [1] start of the last panel, [2] info, [3] reserved
  int n, m, nf, ld;  // n = pivot candidates (rows [0,n)); for a block root n = nT + delayed columns, set on the device
  int nb, pad;        // first border row (static): rows [n, nb) are unused delayed-pivot slots, rows [nb, nf) the border
};

constexpr int ST_KCUR = 0, ST_KPREV = 1, ST_INFO = 2;
constexpr int NBMAX = 64;        // widest panel (columns of W)
constexpr double BK_ALPHA = 0.6403882032022076;  // (1 + sqrt(17)) / 8

}  // namespace ppb
