/*
 * parapint_b200.h -- C ABI of the B200-native Schur-complement KKT solver.
 *
 * The reference (sandialabs/parapint) has no FFI of its own: its hot path is Python calling
 * SciPy/MA27/MUMPS.  These entry points are what a binding for that path binds instead; each cites
 * the reference call site it replaces (paths relative to the reference root).  INTEGRATION.md shows
 * the ctypes stub a parapint maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.  One handle per process / per GPU.
 *  - every function returns a status code equal to parapint's LinearSolverStatus value
 *    (parapint/linalg/results.py:4-9): 0 successful, 1 not_enough_memory, 2 singular, 3 error,
 *    4 warning.  Nothing throws.  pp_last_error() gives a message for code 3 and for PP_MISUSE (-1, see below).
 *  - "local blocks" are the diagonal blocks K_i owned by this rank
 *    (parapint/linalg/schur_complement/mpi_explicit_schur_complement.py:198-203); the coupling
 *    matrix Q / S is replicated on every rank (:142-144, :352-360).
 *  - value / vector buffers may live in host or device memory (`on_device` flag).  Host buffers
 *    are staged through pinned memory owned by the handle; device buffers are used in place.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All work is enqueued on
 *    it; functions that return a status or integers synchronise that stream before returning.
 */
#ifndef PARAPINT_B200_H
#define PARAPINT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pp_handle pp_handle;

enum {
  PP_SUCCESSFUL = 0,
  PP_NOT_ENOUGH_MEMORY = 1,
  PP_SINGULAR = 2,
  PP_ERROR = 3,
  PP_WARNING = 4,
  /* not a LinearSolverStatus: the CALL was wrong (null pointer, call order, malformed description).  Nothing was
   * enqueued; pp_last_error() says what.  A binding raises on it at once -- the same call is wrong on every rank --
   * whereas PP_ERROR (a run-time failure of this rank, e.g. a CUDA error) is a status the ranks must first agree on. */
  PP_MISUSE = -1
};

/* Library / build identification (ABI version, compiled arch). */
int pp_abi_version(void);
const char *pp_build_info(void);
const char *pp_last_error(void);

/* Create / destroy a solver bound to CUDA device `device`. */
int pp_create(int device, pp_handle **out);
int pp_destroy(pp_handle *h);

/* Tunables: "pivot_tol" (absolute zero-pivot tolerance), "panel_width" (dense panel, <= 64),
 * "sparse" (0/1: multifrontal subtree path), "pivot_threshold" (u of the threshold test in subtree
 * fronts, default 0.01), "ordering" (0 auto, 1 minimum degree, 2 nested dissection), "nd_leaf",
 * "sparse_fmax", "sparse_dmax", "sparse_dslot", "sparse_min_n", "pair_weak" (0/1: 2x2 pivot pre-selection from
 * the values hint), "cluster_panel" (0/1: thread-block-cluster panel kernel for tall fronts), "cluster_size" (0 = automatic, or 1/2/4/8
 * CTAs per front), "panel_onchip" (0/1: cluster panel kernel that keeps the panel's rows of L in registers / shared
 * memory), "small_front"
 * (0/1: whole-front shared-memory factorisation when every front of a batch has <= 164 rows), "defer_status"
 * (0/1, single-rank use: pp_numeric_local returns a provisional PP_SUCCESSFUL without synchronising and
 * pp_numeric_coupling -- which must then be given the local Schur buffer -- reports the status of both phases and
 * caches the inertia, one host synchronisation per factorisation), "overlap_groups" (0/1: large dense fronts are
 * factorised as two groups on two streams so that the latency-bound panels of one overlap the updates of the other),
 * "subtree_cluster" (0 = automatic, or 1/2/4/8 CTAs per block for the subtree kernels), "auto_residual" (0/1, single-rank use:
 * pp_solve_backward with host outputs also forms the residual norms, which pp_residual_norms then returns without
 * launching or waiting), "no_fallback", "profile". */
int pp_set_option(pp_handle *h, const char *name, double value);

/*
 * Symbolic phase.  Replaces SchurComplementLinearSolver.do_symbolic_factorization
 * (linalg/schur_complement/explicit_schur_complement.py:44-78) and the structure discovery of
 * MPISchurComplementLinearSolver._get_sc_structure (mpi_explicit_schur_complement.py:228-255).
 *
 *   n_local            number of diagonal blocks owned by this rank
 *   block_n[i]         order n_i of local block i
 *   border_ptr[i..i+1] range in border_rows of the nonzero rows of A_i (ascending coupling-row
 *                      indices; `_BorderMatrix.nonzero_rows`, mpi_explicit_schur_complement.py:43-48)
 *   m_c                coupling dimension (order of Q and S)
 *   nvals              number of numeric input values per factorisation: the concatenation, in a
 *                      caller-chosen fixed order, of the COO data of every local K_i, every local
 *                      A_i and Q (duplicates and explicit zeros allowed,
 *                      interfaces/interface.py:454-456,468-471)
 *   dest_front[k]      destination of value k: local block index 0..n_local-1, n_local for the
 *                      coupling matrix Q, or -1 to ignore the value (e.g. the strict upper triangle:
 *                      like the MA27/MUMPS leaves only the lower triangle is read,
 *                      linalg/mumps_interface.py:51,80, linalg/ma27_interface.py:114-116)
 *   dest_row/dest_col  position inside that front (row >= col).  A front is the symmetric matrix
 *                      [[K_i, .],[A_i(nonzero rows), 0]] of order n_i + m_i; border entry
 *                      (nonzero row a, column j) therefore has row = n_i + a, col = j.
 *   values_hint        optional (may be NULL): representative numeric values of the nvals inputs (host).
 *                      Only the ORDERING uses them: columns whose diagonal is numerically zero (constraint
 *                      multipliers of a KKT block) are paired with a neighbour so that they can be
 *                      eliminated as 2x2 pivots instead of being delayed.  The reference calls the symbolic
 *                      phase with the real KKT matrix (algorithms/interior_point.py:542-552), so values are at
 *                      hand; an all-zero hint (linalg/tests/test_linear_solvers.py:66-71) is equivalent to NULL.
 */
int pp_symbolic(pp_handle *h, int32_t n_local, const int32_t *block_n, const int64_t *border_ptr,
                const int32_t *border_rows, int32_t m_c, int64_t nvals, const int32_t *dest_front,
                const int32_t *dest_row, const int32_t *dest_col, const double *values_hint);

/*
 * Sparse coupling systems.  When the pattern of S = Q - sum_i A_i K_i^-1 A_i^T -- the union over ALL blocks of
 * nonzero_rows(A_i) x nonzero_rows(A_i) and of the pattern of Q, as MPISchurComplementLinearSolver._get_sc_structure
 * builds it (mpi_explicit_schur_complement.py:88-125,228-255) -- is sparse and m_c is not small (the block-tridiagonal
 * S of time-decomposed problems, interfaces/schur_complement/sc_ip_interface.py:274-357), pp_symbolic keeps S as the
 * values of that pattern: the Schur buffers of pp_numeric_local / pp_numeric_coupling then hold pp_schur_size(h)
 * doubles (+ PP_SCHUR_TAIL) instead of m_c * m_c, which is what the caller all-reduces (sc_nnz values, :343), and S
 * is factorised level by level as a block-bordered matrix of its own (block cyclic reduction on a chain).
 *
 * pp_set_coupling_cliques: with several ranks the pattern of S depends on the borders of the blocks of EVERY rank
 * (the reference all-gathers them, :244-247).  Call it before pp_symbolic with the nonzero border rows (ascending) of
 * all blocks of all ranks, in any order that is the same on every rank; without it only the local blocks are used
 * (correct on one rank).  n_cliques = 0 forgets a previous list.
 * pp_coupling_stats: out = {child levels (0 = dense coupling front), pp_schur_size, m_c, order of the dense coupling
 * front at the bottom of the chain, diagonal blocks over all levels, largest front of the levels, blocks of level 1, 0}.
 * Options: "coupling_min_sparse" (smallest m_c held sparse, default 384), "coupling_max_density" (default 0.30).
 */
int pp_set_coupling_cliques(pp_handle *h, int32_t n_cliques, const int64_t *ptr, const int32_t *rows);
int64_t pp_schur_size(const pp_handle *h);
int pp_coupling_stats(pp_handle *h, int64_t out[8]);

/*
 * Device-side regularisation (inertia correction without re-uploading the matrix).  parapint's inertia-correction
 * loop (algorithms/interior_point.py:369-395) refactorises up to ~17 times per iteration with diagonal shifts applied
 * by the problem interface (interfaces/interface.py:590-619; sc_ip_interface.py:903-933,1736-1757): +delta added to
 * the Hessian diagonal of every block (accumulating over the retries of one iteration), -delta SET on the
 * (2,2)/(3,3)/linking-multiplier diagonals, +delta SET on the diagonal of the coupling variables.
 * pp_set_diagonal_classes (before pp_symbolic) declares the class of every local row (concatenated blocks) and
 * coupling row: 0 none, 1..PP_SHIFT_SLOTS the shift slot whose value is ADDED to that diagonal entry; the symbolic
 * phase then reserves those diagonal entries (so a Hessian shift never changes the analysed pattern).
 * pp_set_shifts sets the slot values used by the following factorisations (zero after the classes are set), and
 * pp_numeric_local(h, NULL, PP_VALUES_REUSE, ...) factorises the values of the previous call again -- no gather,
 * no host-to-device transfer, no symbolic phase.  pp_value_uploads counts the transfers of the value array.
 */
#define PP_SHIFT_SLOTS 3
#define PP_VALUES_REUSE 2
int pp_set_diagonal_classes(pp_handle *h, int64_t n_local_rows, const int8_t *cls_local, int32_t m_c,
                            const int8_t *cls_c);
int pp_set_shifts(pp_handle *h, const double *shifts);
int64_t pp_value_uploads(const pp_handle *h);

/*
 * Numeric phase, local part.  Replaces the per-block leaf factorisations and the Schur formation
 * loop (explicit_schur_complement.py:99-121; mpi_explicit_schur_complement.py:292-333): assembles
 * the fronts from `values`, runs the batched Bunch-Kaufman LDL^T of every local front and writes
 * this rank's contribution  -sum_i A_i K_i^{-1} A_i^T  (dense: m_c x m_c, column-major, lower
 * triangle valid; sparse coupling system: the values of the pattern) to `schur_local_dev` (DEVICE
 * pointer, caller owned, so that the caller can SUM-reduce it across ranks as
 * mpi_explicit_schur_complement.py:343 does).  The buffer holds pp_schur_size(h) + PP_SCHUR_TAIL doubles: the tail carries {1 if a local block is singular, 0, n_pos, n_neg,
 * n_zero, 0, 0, 0} of this rank's blocks, so the same all-reduce also agrees the status and sums the
 * inertia across ranks (roles of mpi_explicit_schur_complement.py:21 and :427-429).
 * Returns 0, 2 (a local block is singular) or 3.
 */
#define PP_SCHUR_TAIL 8
int pp_numeric_local(pp_handle *h, const double *values, int on_device, double *schur_local_dev,
                     void *stream);

/*
 * Numeric phase, coupling part.  Replaces "S = Q - sum" and the coupling factorisation
 * (explicit_schur_complement.py:108,122-128; mpi_explicit_schur_complement.py:347-358).
 * `schur_sum_dev` is the (all-reduced) sum of the ranks' contributions (DEVICE pointer).
 */
int pp_numeric_coupling(pp_handle *h, const double *schur_sum_dev, void *stream);

/*
 * Inertia (n_pos, n_neg, n_zero) from the pivots: local blocks and coupling matrix separately so
 * the caller can SUM the local parts across ranks (explicit_schur_complement.py:157-172;
 * mpi_explicit_schur_complement.py:404-436; consumer: algorithms/interior_point.py:371-382).
 */
int pp_inertia_local(pp_handle *h, int64_t out[3]);
int pp_inertia_coupling(pp_handle *h, int64_t out[3]);

/*
 * Solve, phase 1 (explicit_schur_complement.py:141-145; mpi_explicit_schur_complement.py:381-385):
 * forward substitution on every local front.  `rhs_local` is the concatenation of the local
 * blocks' right-hand sides (sum n_i doubles).  Writes this rank's coupling contribution
 * -sum_i A_i K_i^{-1} r_i  (m_c doubles) to `rc_local_dev` (DEVICE pointer, caller owned).
 */
int pp_solve_forward(pp_handle *h, const double *rhs_local, int on_device, double *rc_local_dev,
                     void *stream);

/*
 * Solve, phases 2+3 (explicit_schur_complement.py:147-153; mpi_explicit_schur_complement.py:386-398):
 * x_c = S^{-1}(rhs_c + rc_sum), then back substitution on every local front.  `rc_sum_dev` is the
 * all-reduced coupling contribution (DEVICE pointer).  `rhs_c`, `x_local` (sum n_i) and `x_c` (m_c)
 * are host or device buffers according to `on_device`.
 */
int pp_solve_backward(pp_handle *h, const double *rc_sum_dev, const double *rhs_c, int on_device,
                      double *x_local, double *x_c, void *stream);

/*
 * Iterative refinement (no reference counterpart: the reference relies on its leaves' accuracy; here one
 * refinement step with the ORIGINAL values restores ||Kx-b||/||b|| <= 1e-10 on ill-conditioned late-IPM
 * systems).  The handle remembers the device copies of the last factorised values, right-hand side and
 * solution.
 *   pp_residual_local : r_loc = b_loc - (K x)_loc on this rank's rows; buf_dev (DEVICE, m_c + 2 doubles) receives
 *                       -sum_i A_i x_i (partial coupling rows) and the partial sums |r_loc|^2, |b_loc|^2.
 *                       The caller SUM-reduces buf_dev across ranks (nothing to do on one rank).
 *   pp_residual_norms : r_c = b_c - Q x_c + buf_sum[0:m_c]; out = { |r|^2, |b|^2 } of the whole system.
 *   pp_refine_forward / pp_refine_backward : solve K d = r exactly like pp_solve_forward/backward (the caller
 *                       all-reduces rc in between) and add d to the solution (device; copied out if on_device = 0).
 */
int pp_residual_local(pp_handle *h, double *buf_dev, void *stream);
int pp_residual_norms(pp_handle *h, const double *buf_sum_dev, double out[2], void *stream);
int pp_refine_forward(pp_handle *h, double *rc_local_dev, void *stream);
int pp_refine_backward(pp_handle *h, const double *rc_sum_dev, int on_device, double *x_local, double *x_c,
                       void *stream);

/*
 * With defer_status = 2 pp_numeric_coupling reads the reduced tail of the Schur buffer itself and returns
 * PP_SINGULAR if any rank's local phase was singular, PP_NOT_ENOUGH_MEMORY if any rank's sparse path ran out of
 * delayed-pivot capacity (every rank then repeats pp_numeric_local with defer_status = 0 -- the overflowing rank
 * re-analyses densely -- reduces again and calls pp_numeric_coupling again; the reference's contract for that status
 * is the same: enlarge and retry, interior_point.py:645-651).  pp_schur_tail returns that tail:
 * [singular?, overflow?, n_pos, n_neg, n_zero of all ranks' blocks, 0, 0, 0].
 */
int pp_schur_tail(pp_handle *h, double out[8]);

/* Sizes and introspection. */
int64_t pp_factor_bytes(const pp_handle *h);  /* device bytes held by factors + workspaces */
int64_t pp_local_dim(const pp_handle *h);     /* sum of n_i over local blocks */
int64_t pp_kernel_launches(const pp_handle *h); /* kernels launched by this handle so far */

/*
 * Host-side helper (no CUDA): copy `nseg` byte segments between scattered host arrays and one staging buffer
 * (normally the pinned buffer handed to pp_numeric_local / pp_solve_forward) with up to `threads` worker threads.
 * Segment k is len[k] bytes at ptr[k] <-> staging + off[k]; to_staging = 1 gathers, 0 scatters.  Replaces the
 * per-block Python copies the reference does when it hands each K_i to its leaf (explicit_schur_complement.py:
 * 96-101) -- here all blocks go to the device in one transfer.
 */
int pp_host_copy(int64_t nseg, void *const *ptr, const int64_t *off, const int64_t *len, void *staging,
                 int to_staging, int threads);

/*
 * Host-side helper (no CUDA): *equal = 1 when the bytes of a[k] and b[k] (len[k] each) agree for every k.  An
 * interface that rebuilds its KKT matrix every iteration (interfaces/interface.py:432-491) hands over fresh index
 * arrays; they are compared with the analysed pattern by up to `threads` workers in one call instead of one numpy
 * comparison per leaf (the check mumps_interface.py:82-83 makes before re-using its analysis).
 */
int pp_host_equal(int64_t nseg, void *const *a, void *const *b, const int64_t *len, int threads, int *equal);

/*
 * Host-side helper (no CUDA): announce that a pp_host_copy / pp_stage_values / pp_host_equal call follows shortly --
 * typically issued just before the caller blocks on the GPU (the scatter of the solution follows the wait) or starts
 * walking the matrix whose values it is about to gather.  The pool's workers (up to `threads`) wake up now and spin for
 * the job for at most `spin_ns` nanoseconds (capped at 2 ms) instead of being woken by it, which costs tens of
 * microseconds: as much as the copy of a megabyte.  Without a following job they go back to sleep.  No counterpart in
 * the reference (its per-block copies are single-threaded NumPy, explicit_schur_complement.py:96-101).
 */
int pp_host_wake(int threads, int64_t spin_ns);

/*
 * pp_host_copy + the host-to-device transfer of the values, pipelined: the segments (which must tile the `nvals`
 * doubles of the analysed pattern in order) are gathered into the pinned `staging` buffer in `chunks` shares and
 * every share is sent to the device as soon as it is complete, so the transfer of one overlaps the gather of the
 * next.  A following pp_numeric_local(h, staging, 0, ...) finds the values already on their way and copies nothing.
 */
int pp_stage_values(pp_handle *h, int64_t nseg, void *const *ptr, const int64_t *off, const int64_t *len,
                    void *staging, int threads, int chunks, void *stream);

/*
 * One-shot all-reduce (SUM) of a small device buffer over peer memory -- the exchange step of the path
 * (mpi_explicit_schur_complement.py:343 for the S values, :387 for the coupling right-hand side) when the payload
 * is latency-bound (20 kB and 400 B at BASELINE config 2).  bufs[q] / signal_pads[q] are rank q's contribution buffer
 * and signal pad as mapped in THIS process (e.g. torch symmetric memory: buffer_ptrs, signal_pad_ptrs); `slot` is
 * the first of `world` 32-bit words of the pads reserved for this channel; `seq` is a sequence number that every
 * rank increases by one per call on the channel (the pads start at zero, the first call uses 1).  One kernel on
 * `stream`: announce, wait for the peers, add the contributions in rank order into out_dev (n doubles, local).
 * Every rank gets bit-identical sums.  The caller must not rewrite a contribution buffer before the NEXT call on the
 * channel has been enqueued on every rank (use two buffers alternately).
 */
int pp_peer_allreduce(int world, int rank, const void *const *bufs, void *const *signal_pads, int slot, uint32_t seq,
                      int64_t n, double *out_dev, void *stream);

/*
 * Per-kernel-class device timing (measurement aid; enabled with pp_set_option(h, "profile", 1)).
 * CUDA events are recorded on the launching stream around every launch of each class; this call
 * waits for them and returns accumulated milliseconds and launch counts per class.
 */
enum {
  PP_PROF_ASSEMBLE = 0, /* arena clear + scatter-add of the input values */
  PP_PROF_PANEL = 1,    /* Bunch-Kaufman panel factorisation */
  PP_PROF_SWAPS = 2,    /* row interchanges left of the panel */
  PP_PROF_UPDATE = 3,   /* DMMA trailing update */
  PP_PROF_SCHUR = 4,    /* gather of the local Schur contribution */
  PP_PROF_FORWARD = 5,  /* forward sweep on the local fronts */
  PP_PROF_BACKWARD = 6, /* backward sweep on the local fronts */
  PP_PROF_SUBTREE = 7,  /* multifrontal factorisation of the subtree fronts (one CTA per block) */
  PP_PROF_CLASSES = 8
};
int pp_profile(pp_handle *h, double *ms, int64_t *launches, int reset);

/*
 * Symbolic statistics of local block `block`: out = {plan id, subtree supernodes, root columns,
 * delayed-pivot slots, static nnz(L) of the subtree part, largest subtree front, factor doubles,
 * stack doubles, number of distinct plans, 1 if the sparse path overflowed and the handle fell back
 * to dense fronts, delayed pivots that reached the root in the last factorisation, failure flag}.
 */
int pp_plan_stats(pp_handle *h, int32_t block, int64_t out[12]);

/*
 * Host-only access to the symbolic analysis of ONE block (no GPU needed): minimum-degree ordering,
 * supernodes, subtree / root split and the destination of every input entry.  `rows`/`cols` are
 * lower-triangular positions in the front [K | border rows] (row >= n: border row `row - n`).
 * Arrays: rootcols, col_ptr, cols, row_ptr, rows, rel, parent, nchild, dcap, ent_ptr, tgt_row,
 * tgt_col, tgt_src_ptr, tgt_src, root_row, root_col, root_src.  Scalars: n, m, nT, DR, ns, nnz_l,
 * max_front, l_total, stack_cap.  Pass fmax / dmax / min_sparse_n < 0 for the defaults.
 */
typedef struct pp_plan pp_plan;
int pp_plan_create(int32_t n, int32_t m, int64_t nent, const int32_t *rows, const int32_t *cols,
                   int32_t fmax, int32_t dmax, int32_t min_sparse_n, pp_plan **out);
int pp_plan_set_ordering(int32_t ordering); /* for pp_plan_create: 0 auto, 1 minimum degree, 2 nested dissection */
/* Representative values of the entries of the NEXT pp_plan_create (the role `values_hint` plays in pp_symbolic: columns
 * whose diagonal is numerically zero are ordered as 2x2 pivots with a partner); consumed by that call and ignored when
 * `nent` is not its entry count. */
int pp_plan_set_hint(const double *values, int64_t nent);
int pp_plan_get(const pp_plan *plan, const char *name, const int32_t **data, int64_t *len);
int pp_plan_scalar(const pp_plan *plan, const char *name, int64_t *value);
int pp_plan_destroy(pp_plan *plan);

/*
 * Host-only access to one level of the sparse-coupling analysis (no GPU needed): pattern of S from the cliques and
 * Q, and its view as a block-bordered matrix (blocks = an independent set of variable groups).  Arrays (as int64):
 * scalars = {sparse?, n_blocks, m_next, nnz, m_c}, colptr, rowidx, block_n, border_ptr, border_rows, dest_front,
 * dest_row, dest_col (one per pattern entry), perm_local, perm_c.  min_mc / max_density < 0: defaults.
 */
typedef struct pp_cplan pp_cplan;
int pp_cplan_create(int32_t m_c, int32_t n_cliques, const int64_t *ptr, const int32_t *rows, int64_t nq,
                    const int32_t *qrow, const int32_t *qcol, int32_t min_mc, double max_density, pp_cplan **out);
int pp_cplan_get(const pp_cplan *plan, const char *name, int64_t *buf, int64_t cap, int64_t *len);
int pp_cplan_destroy(pp_cplan *plan);

/* Debug / test access: copy front `f` (0..n_local-1 local, n_local = coupling) to host,
 * column-major with leading dimension *ld; piv/bsz receive the pivot records (n entries). */
int pp_debug_front(pp_handle *h, int32_t f, double *out, int64_t out_len, int32_t *ld, int32_t *piv,
                   int32_t *bsz);

/*
 * Interior-point vector kernels: the O(n) passes parapint's IPM loop makes over its iterates
 * between two linear solves -- fraction_to_the_boundary (algorithms/interior_point.py:655-758), the bound-multiplier
 * steps (interfaces/interface.py:548-570), the complementarity / scaling terms of check_convergence (:241-251,
 * :274-315), the infeasibility maxima (:253-269) and the step update (:587-626) -- each ONE streaming pass over
 * device-resident vectors with the reduction finished on the device.  Stateless (no handle): every pointer is a
 * device pointer; `workspace` is pp_ipm_workspace_bytes() of device memory, zero-filled once by the caller and then
 * only passed along (calls that share a workspace must be ordered on one stream); results are accumulated into small
 * device buffers the caller seeds with pp_ipm_fill and reads back when it needs them.  The arithmetic follows the
 * reference's NumPy expressions operation by operation (no fused multiply-add), so minima and maxima are bit-identical
 * to the reference; sums are formed in a fixed order.  "One group" = the primals with primals_lb / primals_ub and
 * their multipliers, or the slacks with ineq_lb / ineq_ub and theirs; infinite bounds behave as in the reference.
 *
 * pp_ipm_fraction_to_boundary: out2[0] = min(out2[0], largest alpha with x + alpha dx inside the tau-shrunk bounds),
 *     out2[1] = min(out2[1], the same for zl, zu against 0 with their steps formed on the fly).  Seed out2 with 1.
 * pp_ipm_complementarity: out6[0] = max(out6[0], max |(x - lb) zl - barrier| over lb > -inf), out6[1] likewise for ub,
 *     out6[2] += sum |zl|, out6[3] += sum |zu|, out6[4] += #finite lb, out6[5] += #finite ub.  Seed with 0.
 * pp_ipm_max_abs: out2[0] = max(out2[0], max |a - b|) (b may be NULL: max |a|), out2[1] += sum |a|.  Seed with 0.
 * pp_ipm_step: the group's update with alpha3 = [alpha_primal, alpha_dual, line-search step] READ FROM DEVICE MEMORY
 *     (normally what pp_ipm_fraction_to_boundary left there): multiplier steps from the values before the update,
 *     then x += ls (alpha_p dx), zl += ls (alpha_d dzl), zu += ls (alpha_d dzu).
 * pp_ipm_axpy: y += ls (alpha3[which] dy) for the equality / inequality multipliers (which = 1) or any primal-like
 *     vector (which = 0).
 */
int64_t pp_ipm_workspace_bytes(void);
int pp_ipm_fill(double *out, int32_t n, double value, void *stream);
int pp_ipm_fraction_to_boundary(int64_t n, double tau, double barrier, const double *x, const double *dx,
                                const double *lb, const double *ub, const double *zl, const double *zu,
                                double *out2, void *workspace, void *stream);
int pp_ipm_complementarity(int64_t n, double barrier, const double *x, const double *lb, const double *ub,
                           const double *zl, const double *zu, double *out6, void *workspace, void *stream);
int pp_ipm_max_abs(int64_t n, const double *a, const double *b, double *out2, void *workspace, void *stream);
int pp_ipm_step(int64_t n, const double *alpha3, double barrier, double *x, const double *dx, const double *lb,
                const double *ub, double *zl, double *zu, void *stream);
int pp_ipm_axpy(int64_t n, const double *alpha3, int32_t which, double *y, const double *dy, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PARAPINT_B200_H */
