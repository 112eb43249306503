// C ABI of the B200-native Schur-complement KKT solver (see include/parapint_b200.h).
//
// Build (done by __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
//        -I include -o parapint_b200/csrc/libparapint_b200.so parapint_b200/csrc/capi.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/parapint_b200.h"
#include "factor.cuh"
#include "front.cuh"
#include "solve.cuh"
#include "sparse.cuh"
#include "small.cuh"
#include "symbolic.hpp"
#include "hostcopy.hpp"
#include "coupling.hpp"
#include "peer.cuh"
#include "ipmvec.cuh"
#include <map>

using namespace ppb;

namespace {

thread_local std::string g_error;
size_t h_optin_smem = 227 * 1024;  // opt-in shared memory per block of the device (set by pp_create)

struct CudaFail {
  std::string msg;
};

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      throw CudaFail{std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +  \
                     std::to_string(__LINE__) + ")"};                                              \
  } while (0)

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    n = count;
    if (count) CK(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const std::vector<T> &h) {
    alloc(h.size());
    if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

template <class T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0;
  void ensure(size_t count) {
    if (count <= n) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    CK(cudaMallocHost(&p, count * sizeof(T)));
    n = count;  // only after the allocation succeeded
  }
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
};

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// Entry points run on the handle's device and leave the caller's current device as they found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) CK(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

struct pp_handle {
  int device = 0;
  int n_local = 0, m_c = 0;
  bool have_symbolic = false, local_factored = false, coupling_factored = false, forward_done = false;
  std::vector<int> n, m, nf, ld;  // n_local + 1 fronts (last = coupling)
  std::vector<int> nmin;          // pivots every factorisation is sure to eliminate in the front (static columns)
  int64_t local_dim = 0;
  int nmax_local = 0, nfmax_local = 0;
  std::vector<int> block_n;       // original order of every local block
  // options
  double pivot_tol = 0.0;
  int panel_width = 64;
  double pivot_threshold = 0.01;  // u of the threshold test in the subtree fronts
  bool use_sparse = true;
  bool no_fallback = false;
  bool use_cluster = true;
  bool panel_onchip = true;       // cluster panel kernel with the panel's rows of L in registers / shared memory
  bool panel_spec = true;         // speculative panel: all diagonals assumed to pass, one reduction per panel
  int cluster_size = 0;           // 0 = automatic; 1, 2, 4, 8 force the CTAs per front of the cluster panel kernel
  int overlap_groups = 2;         // groups of fronts on separate streams: panels of one overlap updates of the others
  int update_strip = -1;          // tiles per CTA of the pipelined trailing update: -1 automatic, 0/1 one-tile kernel, 2..8
  std::vector<cudaStream_t> aux_streams;
  std::vector<cudaEvent_t> aux_events;
  cudaEvent_t ev_fork = nullptr;
  int subtree_cluster = 0;        // 0 = automatic; 1, 2, 4, 8 force the CTAs per block of the subtree kernels
  int sm_count = 148;
  int defer_status = 0;           // 1 single rank: one host sync per factorisation (status + inertia read together);
                                  // 2 several ranks: pp_numeric_local does not synchronise, its status and overflow
                                  //   flag travel in the tail of the Schur buffer the caller all-reduces
  bool status_pending = false, inertia_cached = false;
  unsigned long long inertia_cache[6] = {0, 0, 0, 0, 0, 0};
  double *last_schur = nullptr;
  bool use_small = true;          // whole-front shared-memory factorisation when every front of a batch fits
  bool sparse_failed = false;     // a block overflowed its delayed-pivot capacity: all blocks were redone dense
  PlanOptions plan_opt;
  // saved symbolic inputs (for the dense re-analysis after a sparse-path overflow)
  std::vector<int32_t> in_block_n, in_border_rows, in_dest_front, in_dest_row, in_dest_col;
  std::vector<int64_t> in_border_ptr;
  std::vector<double> in_hint;    // representative values (ordering heuristics only)
  // multifrontal (subtree) part
  std::vector<PatternPlan> plans;
  std::vector<int> block_plan;
  DevBuf<int> planI;              // all integer tables of all plans
  DevBuf<SnHead> planH;
  DevBuf<int2> planT;
  DevBuf<PlanDev> plans_dev;
  DevBuf<SparseBlock> blocks_dev;
  DevBuf<double> arenaL, arenaStack, ywork, root_rhs, root_x;
  DevBuf<int> arenaBI;
  DevBuf<long long> vec_off, root_off;
  int64_t root_total = 0;
  int max_leaves = 0;             // most level-0 small fronts in any local block
  int leaf_cap = SF_TBUF;         // largest level-0 front (rows) over the local plans
  // iterative refinement: K applied from the input values (row lists), last solve's device vectors
  DevBuf<long long> rl_ptr, rl_src, rb_ptr, rb_src, rq_ptr, rq_src;
  DevBuf<int> rl_col, rb_col, rq_col;
  DevBuf<double> res_loc, res_c, res_part, res_out, dx_tmp, dxc_tmp, bc_keep;
  int res_blocks = 0;
  const double *last_vals = nullptr, *last_rhs = nullptr;
  double *last_x = nullptr, *last_xc = nullptr;
  bool solved = false;
  PinBuf<double> pin_out2;
  PinBuf<double> pin_tail;
  bool tail_valid = false;
  const void *staged_from = nullptr;  // pinned buffer whose contents pp_stage_values already sent to h->vals
  bool auto_residual = false;     // single rank: pp_solve_backward also forms the residual norms (same sync as the copy-out)
  bool norms_valid = false;
  DevBuf<double> res_buf;
  // device storage
  DevBuf<double> arenaA, arenaW, arenaZ, vals, rhs, x, xc, crhs;
  DevBuf<int> arenaI, flag;
  DevBuf<Front> fronts;
  std::vector<Front> hfronts;
  DevBuf<unsigned long long> inertia;  // [0..2] local, [3..5] coupling
  DevBuf<int64_t> asm_dst, asm_ptr, asm_src, src_ptr, brow_ptr, rhs_off, root_off64;
  DevBuf<int32_t> src_front, src_pos, brow;
  DevBuf<int64_t> src_aoff;       // per Schur source: arena offset of (its border row, first border column) in the front
  DevBuf<int32_t> src_ld;         // ... and the front's leading dimension (negative: partial border)
  DevBuf<int64_t> src_boff;       // per Schur source: offset of the front's bvec entry in arenaZ (rc_gather)
  PinBuf<double> pin_vals, pin_vec;
  PinBuf<int> pin_flag;
  PinBuf<unsigned long long> pin_inertia;
  int64_t nvals = 0, nuniq = 0;
  size_t arenaA_elems = 0;
  int64_t launches = 0;
  int64_t bytes = 0;
  // device-side regularisation: class of every local / coupling diagonal entry (0 = none) and the shift slots, which
  // live behind the nvals input values on the device so that a retry of the inertia-correction loop is a refactor only
  std::vector<int8_t> cls_local, cls_c;
  bool have_classes = false;
  double shifts[PP_SHIFT_SLOTS] = {0.0, 0.0, 0.0};
  PinBuf<double> pin_shifts;
  int64_t value_uploads = 0;          // host-to-device transfers of the value array so far (statistics)
  // sparse coupling system (time-decomposed problems): S is kept as the values of its pattern and factorised by a
  // CHILD handle that sees it as a block-bordered matrix once more (coupling.hpp); children are replicated per rank
  CouplingOptions cpl_opt;
  bool have_cliques = false;
  std::vector<int64_t> clq_ptr;       // borders of the blocks of ALL ranks (pp_set_coupling_cliques); else the local ones
  std::vector<int32_t> clq_rows;
  pp_handle *child = nullptr;
  int depth = 0;
  int64_t schur_size = 0;             // doubles of the Schur payload: m_c^2 (dense) or the pattern size (sparse)
  int64_t sc_nnz = 0;
  DevBuf<int32_t> sp_row, sp_col, sp_perm_local, sp_perm_c;
  DevBuf<int64_t> sp_qptr, sp_qsrc;
  DevBuf<double> sp_vals, own_schur, own_rc;
  unsigned long long chain_inertia[3] = {0, 0, 0};
  bool chain_singular = false;
  // optional per-kernel-class timing (option "profile"): CUDA events around every launch
  bool profile = false;
  struct Span { cudaEvent_t a, b; int cls; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[PP_PROF_CLASSES] = {0};
  int64_t prof_launches[PP_PROF_CLASSES] = {0};
  ~pp_handle() {
    for (auto &s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : event_pool) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (auto e : aux_events) cudaEventDestroy(e);
    for (auto a : aux_streams) cudaStreamDestroy(a);
    delete child;
  }
};

namespace {

cudaEvent_t take_event(pp_handle *h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  return e;
}

// RAII span: records an event pair around the launches of one kernel class when profiling is on.
struct ProfSpan {
  pp_handle *h;
  cudaStream_t st;
  pp_handle::Span s;
  bool on;
  ProfSpan(pp_handle *h_, int cls, cudaStream_t st_) : h(h_), st(st_), on(h_->profile) {
    if (!on) return;
    s.cls = cls;
    s.a = take_event(h);
    s.b = take_event(h);
    cudaEventRecord(s.a, st);
  }
  ~ProfSpan() {
    if (!on) return;
    cudaEventRecord(s.b, st);
    h->spans.push_back(s);
  }
};

void resolve_profile(pp_handle *h) {
  for (auto &s : h->spans) {
    float ms = 0.f;
    cudaEventSynchronize(s.b);
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      h->prof_ms[s.cls] += ms;
      h->prof_launches[s.cls] += 1;
    }
    h->event_pool.push_back(s.a);
    h->event_pool.push_back(s.b);
  }
  h->spans.clear();
}

bool is_pinned_host(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// CTAs per block for the subtree kernels.  Measured on B200 with config-2 blocks (tools/cluster_probe.py): the
// kernels are bound by the DEPTH of the assembly tree (8 blocks take 148 us, 64 blocks 176 us), so more CTAs per block
// buy little -- two CTAs shave 5 % off the factorisation (the two levels with ~24 medium fronts need fewer rounds),
// more only add cluster barriers, and the solves (a few microseconds per level) lose to the barriers altogether.
int subtree_csize(const pp_handle *h, bool factor) {
  if (h->subtree_cluster > 0) return h->subtree_cluster;
  return (factor && h->n_local * 2 <= h->sm_count) ? 2 : 1;
}

// launch `kern` with one thread-block cluster of `csize` CTAs per block
template <class... P, class... A>
void launch_clustered(void (*kern)(P...), int nblocks, int csize, int threads, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nblocks * csize));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...));
}

// Factor fronts [first, first+count): fixed schedule of (panel, interchange, update) launches.  A
// panel eliminates NB-1 or NB columns, so after launch `it` front f has done at least
// min(n_f, (it+1)(NB-1)) columns; that bounds the tile grid of the update from the host side.
void factor_group(pp_handle *h, int first, int count, cudaStream_t st);

// The panel kernel is latency-bound (a chain of pivot columns per front), the trailing update throughput-bound, and
// inside ONE front they cannot overlap: Bunch-Kaufman may pick its pivot anywhere in the trailing matrix, so a panel
// needs the whole update before it.  Different fronts are independent, though: the batch is cut into two groups that
// run the same launch sequence on two streams, so one group's panels fill the gaps of the other group's updates.
void factor_fronts(pp_handle *h, int first, int count, cudaStream_t st) {
  if (count == 0) return;
  int nfmax = 0;
  for (int f = first; f < first + count; ++f) nfmax = std::max(nfmax, h->nf[f]);
  const int G = std::min(h->overlap_groups, count);
  if (G < 2 || nfmax < 512) {
    factor_group(h, first, count, st);
    return;
  }
  if (!h->ev_fork) CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  while ((int)h->aux_streams.size() < G - 1) {
    cudaStream_t s2;
    cudaEvent_t e2;
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    h->aux_streams.push_back(s2);
    h->aux_events.push_back(e2);
  }
  CK(cudaEventRecord(h->ev_fork, st));
  for (int g = 1; g < G; ++g) CK(cudaStreamWaitEvent(h->aux_streams[g - 1], h->ev_fork, 0));
  for (int g = 0; g < G; ++g) {
    const int lo = (int)((int64_t)count * g / G), hi = (int)((int64_t)count * (g + 1) / G);
    factor_group(h, first + lo, hi - lo, g == 0 ? st : h->aux_streams[g - 1]);
  }
  for (int g = 1; g < G; ++g) {
    CK(cudaEventRecord(h->aux_events[g - 1], h->aux_streams[g - 1]));
    CK(cudaStreamWaitEvent(st, h->aux_events[g - 1], 0));
  }
}

void factor_group(pp_handle *h, int first, int count, cudaStream_t st) {
  if (count == 0) return;
  const int NB = h->panel_width;
  const Front *fr = h->fronts.p + first;
  int nmax = 0, nfmax = 0;
  for (int f = first; f < first + count; ++f) {
    nmax = std::max(nmax, h->n[f]);
    nfmax = std::max(nfmax, h->nf[f]);
  }
  if (nmax == 0) return;
  if (h->use_small && nfmax <= SM_CAP) {  // nf = static n + m bounds the active rows of every front
    ProfSpan sp(h, PP_PROF_PANEL, st);
    front_small_kernel<<<count, SF_NT, SM_SMEM, st>>>(fr, h->pivot_threshold, h->pivot_tol);
    h->launches++;
    CK(cudaGetLastError());
    return;
  }
  const bool small = nfmax <= 384;  // short columns: four warps per front keep the block reductions cheap
  // tall fronts: a cluster of CTAs per front, as many as fill the GPU once (8 is the portable maximum)
  int csize = 1;
  if (h->use_cluster && nfmax >= 1024)
    while (csize < 8 && count * csize * 2 <= 128 && PC_NT * csize * 2 <= nfmax + PC_NT) csize *= 2;
  if (h->use_cluster && h->panel_onchip && nfmax >= 1024) {
    // on-chip panel: one CTA per SM (196 KB of shared memory), about one row per thread (more CTAs per front than
    // rows to fill them only lengthen the cluster barriers)
    csize = 1;
    while (csize < 8 && count * csize * 2 <= 148 && PC_NT * csize * 2 <= nfmax + PC_NT / 2) csize *= 2;
  }
  if (h->use_cluster && h->cluster_size > 0 && nfmax >= 1024) csize = h->cluster_size;
  const int iters = (nmax + (NB - 2)) / (NB - 1);
  for (int it = 0; it < iters; ++it) {
    {
      ProfSpan sp(h, PP_PROF_PANEL, st);
      if (small) {
        front_panel_kernel<128><<<count, 128, 0, st>>>(fr, NB, h->pivot_tol);
      } else if (csize > 1 || (h->panel_spec && h->panel_onchip && (nfmax >= 1024 || count * 2 <= h->sm_count))) {
        // (the on-chip kernel needs an SM per CTA: many mid-sized fronts are better off several to an SM below)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(count * csize);
        cfg.blockDim = dim3(PC_NT);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (h->panel_onchip) {
          cfg.dynamicSmemBytes = OC_SMEM;
          CK(cudaLaunchKernelEx(&cfg, front_panel_cluster_oc_kernel, fr, NB, h->pivot_tol, h->panel_spec ? 1 : 0));
        } else {
          CK(cudaLaunchKernelEx(&cfg, front_panel_cluster_kernel, fr, NB, h->pivot_tol));
        }
      } else {
        front_panel_kernel<512><<<count, 512, 0, st>>>(fr, NB, h->pivot_tol);
      }
      h->launches++;
    }
    if (it > 0) {
      ProfSpan sp(h, PP_PROF_SWAPS, st);
      dim3 g((nmax + 255) / 256, count);
      front_swaps_left_kernel<<<g, 256, 0, st>>>(fr);
      h->launches++;
    }
    int ntiles = 0;
    int64_t all_tiles = 0;
    for (int f = first; f < first + count; ++f) {
      if (h->n[f] <= it * (NB - 1)) continue;  // finished in an earlier launch
      // lower bound on the columns done: a root eliminates at least its static columns (nmin), the
      // delayed-pivot slots may be unused
      const int done = std::min(h->nmin[f], (it + 1) * (NB - 1));
      if (done >= h->nf[f]) continue;
      const int t = update_tile_count(h->nf[f], done);
      ntiles = std::max(ntiles, t);
      all_tiles += t;
    }
    if (ntiles > 0) {
      ProfSpan sp(h, PP_PROF_UPDATE, st);
      // strips of S tiles per CTA (software-pipelined kernel, one CTA per SM) when strips of at least eight tiles still
      // fill the GPU four times over; otherwise one tile per CTA, two CTAs per SM (measured, tools/update_probe.cu:
      // 32 / 16 / 4 fronts of the config-5 root shape: 0.96 / 0.51 / 0.16 ms with strips against 1.06 / 0.54 / 0.15 ms)
      int S = h->update_strip >= 0 ? h->update_strip : (int)std::min<int64_t>(12, all_tiles / (4 * (int64_t)h->sm_count));
      if (h->update_strip < 0 && S < 8) S = 0;
      S = std::min(S, UPS_MAX);
      if (S >= 2) {
        int nstrips = 0;
        for (int f = first; f < first + count; ++f) {
          if (h->n[f] <= it * (NB - 1)) continue;
          const int done = std::min(h->nmin[f], (it + 1) * (NB - 1));
          if (done >= h->nf[f]) continue;
          nstrips = std::max(nstrips, update_strip_count(h->nf[f], done, S));
        }
        dim3 g(nstrips, count);
        front_update_strip_kernel<<<g, UPD_THREADS, UPS_SMEM, st>>>(fr, S);
      } else {
        dim3 g(ntiles, count);
        front_update_kernel<<<g, UPD_THREADS, UPD_SMEM, st>>>(fr);
      }
      h->launches++;
    }
  }
  CK(cudaGetLastError());
}

// flag[slot] = any front of [first, first+count) reported a zero pivot (stream-ordered, no host sync)
void post_flag(pp_handle *h, int first, int count, int slot, cudaStream_t st) {
  if (count == 0) {
    CK(cudaMemsetAsync(h->flag.p + slot, 0, sizeof(int), st));
    return;
  }
  collect_info_kernel<<<1, 256, 0, st>>>(h->fronts.p + first, count, h->flag.p + slot);
  h->launches++;
}

// flags (8 ints) and inertia (6 counters) to the host with one synchronisation
void fetch_status(pp_handle *h, cudaStream_t st) {
  h->pin_flag.ensure(8);
  h->pin_inertia.ensure(8);
  CK(cudaMemcpyAsync(h->pin_flag.p, h->flag.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h->pin_inertia.p, h->inertia.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
}

int read_flag(pp_handle *h, int first, int count, cudaStream_t st) {
  if (count == 0) return 0;
  post_flag(h, first, count, 0, st);
  h->pin_flag.ensure(8);
  CK(cudaMemcpyAsync(h->pin_flag.p, h->flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return h->pin_flag.p[0];
}

size_t solve_smem(int nf) { return ((size_t)((nf + 1) & ~1) + SB * SPITCH) * sizeof(double); }

template <class F>
int guarded(F &&body) {
  try {
    return body();
  } catch (const CudaFail &e) {
    g_error = e.msg;
    return PP_ERROR;
  } catch (const std::bad_alloc &) {
    g_error = "host allocation failed";
    return PP_NOT_ENOUGH_MEMORY;
  } catch (const std::exception &e) {
    g_error = e.what();
    return PP_ERROR;
  }
}

int fail(const std::string &msg) {
  g_error = msg;
  return PP_ERROR;
}

// API misuse (null pointer, wrong call order, malformed description): not a LinearSolverStatus
int misuse(const std::string &msg) {
  g_error = msg;
  return PP_MISUSE;
}

}  // namespace

extern "C" {

int pp_abi_version(void) { return 1; }

const char *pp_build_info(void) { return "parapint_b200 sm_100a fp64 dmma-m8n8k4 abi1"; }

const char *pp_last_error(void) { return g_error.c_str(); }

int pp_create(int device, pp_handle **out) {
  if (!out) return misuse("pp_create: null out pointer");
  *out = nullptr;
  return guarded([&]() {
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return misuse("pp_create: no such CUDA device");
    DeviceGuard dev_guard(device);
    // Every kernel that uses dynamic shared memory may use up to the device's opt-in limit (227 KB on B200); set
    // once per device -- a per-launch setting would let a child level LOWER the limit its parent still needs.
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if ((size_t)optin < SM_SMEM) return fail("pp_create: the device offers too little shared memory per block");
    const auto lim = cudaFuncAttributeMaxDynamicSharedMemorySize;
    CK(cudaFuncSetAttribute(front_update_kernel, lim, (int)UPD_SMEM));
    CK(cudaFuncSetAttribute(front_update_strip_kernel, lim, (int)UPS_SMEM));
    CK(cudaFuncSetAttribute(front_panel_cluster_oc_kernel, lim, (int)OC_SMEM));
    CK(cudaFuncSetAttribute(subtree_factor_kernel, lim, (int)SF_SMEM));
    CK(cudaFuncSetAttribute(front_small_kernel, lim, (int)SM_SMEM));
    CK(cudaFuncSetAttribute(subtree_forward_kernel, lim, (int)SV_SMEM));
    CK(cudaFuncSetAttribute(subtree_backward_kernel, lim, (int)SV_SMEM));
    // (the opt-in limit covers static + dynamic shared memory of a kernel)
    auto allow_all = [&](auto kern) {
      cudaFuncAttributes fa;
      CK(cudaFuncGetAttributes(&fa, kern));
      CK(cudaFuncSetAttribute(kern, lim, optin - (int)fa.sharedSizeBytes));
    };
    allow_all(front_forward_kernel<512>);
    allow_all(front_backward_kernel<512>);
    allow_all(coupling_solve_kernel<512>);
    allow_all(front_forward_cluster_kernel);
    allow_all(front_backward_cluster_kernel);
    allow_all(coupling_solve_cluster_kernel);
    allow_all(subtree_leaf_kernel<8>);
    allow_all(subtree_leaf_kernel<16>);
    allow_all(subtree_leaf_kernel<32>);
    allow_all(subtree_leaf_forward_kernel<8>);
    allow_all(subtree_leaf_forward_kernel<16>);
    allow_all(subtree_leaf_forward_kernel<32>);
    allow_all(subtree_leaf_backward_kernel<8>);
    allow_all(subtree_leaf_backward_kernel<16>);
    allow_all(subtree_leaf_backward_kernel<32>);
    h_optin_smem = (size_t)optin - 1024;
    auto *h = new pp_handle();
    h->device = device;
    CK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    h->flag.alloc(8);
    h->inertia.alloc(8);
    CK(cudaMemset(h->inertia.p, 0, 8 * sizeof(unsigned long long)));
    CK(cudaMemset(h->flag.p, 0, 8 * sizeof(int)));
    *out = h;
    return (int)PP_SUCCESSFUL;
  });
}

int pp_destroy(pp_handle *h) {
  if (!h) return PP_SUCCESSFUL;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  delete h;
  if (prev >= 0) cudaSetDevice(prev);
  return PP_SUCCESSFUL;
}

int pp_set_option(pp_handle *h, const char *name, double value) {
  if (!h || !name) return misuse("pp_set_option: null argument");
  const std::string key(name);
  if (key == "pivot_tol") {
    if (value < 0) return misuse("pivot_tol must be >= 0");
    h->pivot_tol = value;
  } else if (key == "panel_width") {
    const int nb = (int)value;
    if (nb < 4 || nb > NBMAX) return misuse("panel_width must be in [4, 64]");
    h->panel_width = nb;
  } else if (key == "sparse") {
    h->use_sparse = value != 0.0;
  } else if (key == "panel_onchip") {
    h->panel_onchip = value != 0.0;
  } else if (key == "panel_spec") {
    h->panel_spec = value != 0.0;
  } else if (key == "cluster_size") {
    const int c = (int)value;
    if (c != 0 && c != 1 && c != 2 && c != 4 && c != 8) return misuse("cluster_size must be 0, 1, 2, 4 or 8");
    h->cluster_size = c;
  } else if (key == "overlap_groups") {
    h->overlap_groups = std::max(1, std::min((int)value, 8));
  } else if (key == "update_strip") {
    const int c = (int)value;
    if (c < -1 || c > UPS_MAX) return misuse("update_strip must be -1 (automatic) or in [0, 16]");
    h->update_strip = c;
  } else if (key == "subtree_cluster") {
    const int c = (int)value;
    if (c != 0 && c != 1 && c != 2 && c != 4 && c != 8) return misuse("subtree_cluster must be 0, 1, 2, 4 or 8");
    h->subtree_cluster = c;
  } else if (key == "small_front") {
    h->use_small = value != 0.0;
  } else if (key == "auto_residual") {
    h->auto_residual = value != 0.0;
  } else if (key == "defer_status") {
    h->defer_status = (int)value;
  } else if (key == "pivot_threshold") {
    if (!(value > 0.0 && value <= 0.5)) return misuse("pivot_threshold must be in (0, 0.5]");
    h->pivot_threshold = value;
  } else if (key == "cluster_panel") {
    h->use_cluster = value != 0.0;
  } else if (key == "no_fallback") {
    h->no_fallback = value != 0.0;
  } else if (key == "sparse_dslot") {
    h->plan_opt.dslot = std::max(0, std::min((int)value, 32));
  } else if (key == "pair_weak") {
    h->plan_opt.pair_weak = value != 0.0;
  } else if (key == "ordering") {
    h->plan_opt.ordering = (int)value;
  } else if (key == "nd_leaf") {
    h->plan_opt.nd_leaf = std::max(4, (int)value);
  } else if (key == "sparse_fmax") {
    h->plan_opt.fmax = std::max(8, std::min((int)value, SF_SBUF - 8));
  } else if (key == "sparse_dmax") {
    h->plan_opt.dmax = std::max(0, std::min((int)value, SF_SBUF / 2));
  } else if (key == "sparse_min_n") {
    h->plan_opt.min_sparse_n = (int)value;
  } else if (key == "coupling_min_sparse") {
    h->cpl_opt.min_mc = std::max(2, (int)value);
  } else if (key == "coupling_max_density") {
    if (!(value >= 0.0 && value <= 1.0)) return misuse("coupling_max_density must be in [0, 1]");
    h->cpl_opt.max_density = value;
  } else if (key == "profile") {
    h->profile = value != 0.0;
  } else if (key == "use_graph" || key == "refine_steps") {
    // accepted for forward compatibility; no effect in this build
  } else {
    return misuse("pp_set_option: unknown option " + key);
  }
  return PP_SUCCESSFUL;
}

// Builds every device structure from the saved symbolic inputs.  force_dense: no subtree part.
static int do_symbolic(pp_handle *h, bool force_dense) {
  const int n_local = h->n_local, m_c = h->m_c;
  const int64_t nvals = h->nvals;
  const int32_t *block_n = h->in_block_n.data();
  const int64_t *border_ptr = h->in_border_ptr.data();
  const int32_t *border_rows = h->in_border_rows.data();
  const int32_t *dest_front = h->in_dest_front.data();
  const int32_t *dest_row = h->in_dest_row.data();
  const int32_t *dest_col = h->in_dest_col.data();
  h->have_symbolic = h->local_factored = h->coupling_factored = h->forward_done = false;
  const int nfronts = n_local + 1;
  h->block_n.assign(block_n, block_n + n_local);
  h->local_dim = 0;
  std::vector<int64_t> bptr(nfronts + 1, 0);
  std::vector<int> mloc(n_local, 0);
  for (int f = 0; f < n_local; ++f) {
    mloc[f] = (int)(border_ptr[f + 1] - border_ptr[f]);
    h->local_dim += block_n[f];
    bptr[f + 1] = bptr[f] + mloc[f];
  }
  bptr[nfronts] = bptr[n_local];

  // ---- per-block entry lists and pattern de-duplication ----
  std::vector<std::vector<int>> e_row(n_local), e_col(n_local), e_src(n_local);
  std::vector<std::vector<double>> e_val(n_local);
  const bool have_hint = (int64_t)h->in_hint.size() == nvals && nvals > 0;
  std::vector<int64_t> first_k(n_local, -1);
  std::vector<int64_t> coupling_k;
  for (int64_t k = 0; k < nvals; ++k) {
    const int f = dest_front[k];
    if (f < 0) continue;
    if (f > n_local) return misuse("pp_symbolic: dest_front out of range");
    const int r = dest_row[k], c = dest_col[k];
    if (f == n_local) {
      if (r < 0 || c < 0 || r >= m_c || c > r) return misuse("pp_symbolic: coupling entry outside the lower triangle");
      coupling_k.push_back(k);
      continue;
    }
    if (r < 0 || c < 0 || r >= block_n[f] + mloc[f] || c > r || c >= block_n[f])
      return misuse("pp_symbolic: destination outside the lower triangle of its front");
    if (first_k[f] < 0) first_k[f] = k;
    if (k - first_k[f] > 0x7fffffff) return fail("pp_symbolic: block values span more than 2^31 entries");
    e_row[f].push_back(r);
    e_col[f].push_back(c);
    e_src[f].push_back((int)(k - first_k[f]));
    if (have_hint) e_val[f].push_back(h->in_hint[(size_t)k]);
  }
  // diagonal entries that can be shifted on the device: one more (duplicate) source per classified row, a negative
  // source -c standing for shift slot c - 1 (the same encoding in every block, so equal patterns still share a plan)
  if (h->have_classes) {
    if ((int64_t)h->cls_local.size() != h->local_dim || (int)h->cls_c.size() != m_c)
      return misuse("pp_symbolic: the diagonal classes do not match the block sizes (pp_set_diagonal_classes)");
    int64_t off = 0;
    for (int f = 0; f < n_local; ++f) {
      for (int i = 0; i < block_n[f]; ++i) {
        const int c = h->cls_local[(size_t)(off + i)];
        if (c <= 0) continue;
        if (first_k[f] < 0) first_k[f] = 0;
        e_row[f].push_back(i);
        e_col[f].push_back(i);
        e_src[f].push_back(-c);
        if (have_hint) e_val[f].push_back(0.0);
      }
      off += block_n[f];
    }
  }
  h->plans.clear();
  h->block_plan.assign(n_local, -1);
  {
    std::map<std::vector<int>, int> seen;
    for (int f = 0; f < n_local; ++f) {
      std::vector<int> key;
      key.reserve(3 * e_row[f].size() + 2);
      key.push_back(block_n[f]);
      key.push_back(mloc[f]);
      key.insert(key.end(), e_row[f].begin(), e_row[f].end());
      key.insert(key.end(), e_col[f].begin(), e_col[f].end());
      key.insert(key.end(), e_src[f].begin(), e_src[f].end());
      auto it = seen.find(key);
      if (it == seen.end()) {
        const int id = (int)h->plans.size();
        h->plans.push_back(build_plan(block_n[f], mloc[f], e_row[f], e_col[f], e_src[f], h->plan_opt,
                                      force_dense || !h->use_sparse, have_hint ? &e_val[f] : nullptr));
        seen.emplace(std::move(key), id);
        h->block_plan[f] = id;
      } else {
        h->block_plan[f] = it->second;
      }
    }
  }

  // ---- dense fronts: one root per block + the coupling front ----
  h->n.assign(nfronts, 0);
  h->nmin.assign(nfronts, 0);
  h->m.assign(nfronts, 0);
  h->nf.assign(nfronts, 0);
  h->ld.assign(nfronts, 0);
  h->nmax_local = h->nfmax_local = 0;
  for (int f = 0; f < n_local; ++f) {
    const PatternPlan &P = h->plans[h->block_plan[f]];
    h->n[f] = P.nT + P.DR;
    h->nmin[f] = P.nT;
    h->m[f] = mloc[f];
  }
  h->n[n_local] = h->child ? 0 : m_c;  // sparse coupling system: no dense coupling front
  h->nmin[n_local] = h->n[n_local];
  h->m[n_local] = 0;
  std::vector<size_t> offA(nfronts), offW(nfronts), offZ(nfronts), offI(nfronts);
  size_t totA = 0, totW = 0, totZ = 0, totI = 0;
  for (int f = 0; f < nfronts; ++f) {
    h->nf[f] = h->n[f] + h->m[f];
    h->ld[f] = std::max(16, round_up(h->nf[f], 16));
    offA[f] = totA;
    totA += (size_t)h->ld[f] * std::max(h->nf[f], 1);
    offW[f] = totW;
    totW += (size_t)h->ld[f] * NBMAX;
    offZ[f] = totZ;
    totZ += (size_t)round_up(h->n[f] + h->m[f] + 2, 16);
    offI[f] = totI;
    totI += (size_t)3 * round_up(h->n[f] + 1, 16) + 16;
    if (f < n_local) {
      h->nmax_local = std::max(h->nmax_local, h->n[f]);
      h->nfmax_local = std::max(h->nfmax_local, h->nf[f]);
    }
    if (cluster_solve_smem(h->nf[f], CS_MAXC) > h_optin_smem)
      return fail("pp_symbolic: dense front too large for the shared-memory solve vector (more than ~200 000 rows)");
  }
  h->arenaA_elems = totA;
  h->arenaA.alloc(totA);
  h->arenaW.alloc(totW);
  h->arenaZ.alloc(totZ);
  h->arenaI.alloc(totI);
  CK(cudaMemset(h->arenaW.p, 0, totW * sizeof(double)));
  CK(cudaMemset(h->arenaZ.p, 0, totZ * sizeof(double)));
  CK(cudaMemset(h->arenaI.p, 0, totI * sizeof(int)));
  std::vector<Front> hf(nfronts);
  for (int f = 0; f < nfronts; ++f) {
    Front &F = hf[f];
    F.A = h->arenaA.p + offA[f];
    F.W = h->arenaW.p + offW[f];
    F.zbuf = h->arenaZ.p + offZ[f];
    F.bvec = F.zbuf + h->n[f];
    const int seg = round_up(h->n[f] + 1, 16);
    F.ipiv = h->arenaI.p + offI[f];
    F.bsz = F.ipiv + seg;
    F.perm = F.bsz + seg;
    F.state = F.perm + seg;
    F.n = h->n[f];
    F.m = h->m[f];
    F.nf = h->nf[f];
    F.ld = h->ld[f];
    F.nb = h->n[f];
    F.pad = 0;
  }
  h->fronts.upload(hf);
  h->hfronts = hf;

  // ---- dense assembly map (root entries of every block + Q): group values by destination ----
  std::vector<std::pair<int64_t, int64_t>> keyed;
  for (int f = 0; f < n_local; ++f) {
    const PatternPlan &P = h->plans[h->block_plan[f]];
    for (size_t i = 0; i < P.root_row.size(); ++i)
      keyed.emplace_back((int64_t)(offA[f] + (size_t)P.root_row[i] + (size_t)P.root_col[i] * h->ld[f]),
                         P.root_src[i] >= 0 ? first_k[f] + P.root_src[i] : nvals + (-P.root_src[i] - 1));
  }
  if (h->have_classes && !h->child)
    for (int g = 0; g < m_c; ++g)
      if (h->cls_c[(size_t)g] > 0)
        keyed.emplace_back((int64_t)(offA[n_local] + (size_t)g + (size_t)g * h->ld[n_local]), nvals + (h->cls_c[(size_t)g] - 1));
  for (int64_t k : coupling_k)
    if (!h->child) keyed.emplace_back((int64_t)(offA[n_local] + (size_t)dest_row[k] + (size_t)dest_col[k] * h->ld[n_local]), k);
  std::stable_sort(keyed.begin(), keyed.end(), [](const auto &a, const auto &b) {
    return a.first != b.first ? a.first < b.first : a.second < b.second;
  });
  std::vector<int64_t> dst, ptr, src(keyed.size());
  for (size_t i = 0; i < keyed.size(); ++i) {
    if (i == 0 || keyed[i].first != keyed[i - 1].first) {
      dst.push_back(keyed[i].first);
      ptr.push_back((int64_t)i);
    }
    src[i] = keyed[i].second;
  }
  ptr.push_back((int64_t)keyed.size());
  h->nuniq = (int64_t)dst.size();
  h->asm_dst.upload(dst);
  h->asm_ptr.upload(ptr);
  h->asm_src.upload(src);
  if (h->vals.n < (size_t)(nvals + PP_SHIFT_SLOTS)) h->vals.alloc((size_t)(nvals + PP_SHIFT_SLOTS));
  CK(cudaMemset(h->vals.p + nvals, 0, PP_SHIFT_SLOTS * sizeof(double)));

  // ---- plans on the device: supernode headers, packed entry targets, one int table ----
  std::vector<int> ti;
  std::vector<SnHead> heads;
  std::vector<int2> tgts;
  struct Off { size_t heads, tgt, rootcols, cols, rows, rel, child_idx, root_children, tiny_ptr, tiny_idx, med_ptr, med_idx, big_ptr, big_idx, tgt_src; };
  std::vector<Off> offs(h->plans.size());
  auto put = [&](const std::vector<int> &v) { const size_t o = ti.size(); ti.insert(ti.end(), v.begin(), v.end()); ti.push_back(0); return o; };
  for (size_t q = 0; q < h->plans.size(); ++q) {
    const PatternPlan &P = h->plans[q];
    Off &o = offs[q];
    o.rootcols = put(P.rootcols); o.cols = put(P.cols); o.rows = put(P.rows); o.rel = put(P.rel);
    o.child_idx = put(P.child_idx); o.root_children = put(P.root_children); o.tiny_ptr = put(P.tiny_ptr);
    o.tiny_idx = put(P.tiny_idx); o.med_ptr = put(P.med_ptr); o.med_idx = put(P.med_idx); o.big_ptr = put(P.big_ptr);
    o.big_idx = put(P.big_idx); o.tgt_src = put(P.tgt_src);
    o.heads = heads.size();
    o.tgt = tgts.size();
    for (int sidx = 0; sidx < P.ns; ++sidx) {
      SnHead H;
      H.c0 = P.col_ptr[sidx]; H.nc = P.col_ptr[sidx + 1] - P.col_ptr[sidx];
      H.r0 = P.row_ptr[sidx]; H.ncb = P.row_ptr[sidx + 1] - P.row_ptr[sidx];
      H.ch0 = P.child_ptr[sidx]; H.nch = P.child_ptr[sidx + 1] - P.child_ptr[sidx];
      H.dcap = P.dcap[sidx]; H.dslot = P.dslot[sidx];
      H.fid_off = P.fid_off[sidx]; H.fs_off = P.fs_off[sidx]; H.vec_off = P.vec_off[sidx];
      H.ent0 = P.ent_ptr[sidx]; H.nent = P.ent_ptr[sidx + 1] - P.ent_ptr[sidx];
      H.parent = P.parent[sidx]; H.pad0 = H.pad1 = 0;
      H.l_off = P.l_off[sidx]; H.cb_off = P.cb_off[sidx];
      heads.push_back(H);
    }
    for (size_t e = 0; e < P.tgt_row.size(); ++e) {
      const int cntv = P.tgt_src_ptr[e + 1] - P.tgt_src_ptr[e];
      if (cntv > 32767 || P.tgt_row[e] > 255 || P.tgt_col[e] > 255) return fail("pp_symbolic: entry target does not fit its packed record");
      int2 t;
      t.x = P.tgt_row[e] | (P.tgt_col[e] << 8) | (cntv << 16);
      t.y = cntv == 1 ? P.tgt_src[P.tgt_src_ptr[e]] : P.tgt_src_ptr[e];
      tgts.push_back(t);
    }
  }
  if (heads.empty()) heads.push_back(SnHead{});
  if (tgts.empty()) tgts.push_back(int2{0, 0});
  h->planI.upload(ti);
  h->planH.upload(heads);
  h->planT.upload(tgts);
  std::vector<PlanDev> pd(h->plans.size());
  for (size_t q = 0; q < h->plans.size(); ++q) {
    const PatternPlan &P = h->plans[q];
    const Off &o = offs[q];
    PlanDev &D = pd[q];
    const int *b = h->planI.p;
    D.n = P.n; D.m = P.m; D.nT = P.nT; D.DR = P.DR; D.ns = P.ns; D.nlevels = P.nlevels;
    D.nrootch = (int)P.root_children.size(); D.pad = 0;
    D.heads = h->planH.p + o.heads;
    D.tgt = h->planT.p + o.tgt;
    D.rootcols = b + o.rootcols; D.cols = b + o.cols; D.rows = b + o.rows; D.rel = b + o.rel;
    D.child_idx = b + o.child_idx; D.root_children = b + o.root_children; D.tiny_ptr = b + o.tiny_ptr;
    D.tiny_idx = b + o.tiny_idx; D.med_ptr = b + o.med_ptr; D.med_idx = b + o.med_idx; D.big_ptr = b + o.big_ptr;
    D.big_idx = b + o.big_idx; D.tgt_src = b + o.tgt_src;
  }
  h->plans_dev.upload(pd);

  // ---- per-block storage of the subtree part ----
  size_t totL = 0, totS = 0, totBI = 0;
  std::vector<size_t> oL(n_local), oS(n_local), oBI(n_local);
  std::vector<long long> voff(n_local + 1, 0), roff_root(n_local + 1, 0);
  for (int f = 0; f < n_local; ++f) {
    const PatternPlan &P = h->plans[h->block_plan[f]];
    oL[f] = totL; totL += (size_t)P.l_total + 8;
    oS[f] = totS; totS += (size_t)P.cb_total + (size_t)P.vec_total + 16;
    oBI[f] = totBI; totBI += (size_t)P.fid_total + 2 * (size_t)P.fs_total + 3 * (size_t)P.ns + P.nT + P.DR + 16;
    voff[f + 1] = voff[f] + block_n[f];
    roff_root[f + 1] = roff_root[f] + P.nT + P.DR;
  }
  h->root_total = roff_root[n_local];
  h->max_leaves = 0;
  for (int f = 0; f < n_local; ++f) {
    const PatternPlan &P = h->plans[h->block_plan[f]];
    if (P.nlevels > 0) h->max_leaves = std::max(h->max_leaves, P.tiny_ptr[1] - P.tiny_ptr[0]);
  }
  h->leaf_cap = 4;
  for (const PatternPlan &P : h->plans)
    for (int k = P.nlevels > 0 ? P.tiny_ptr[0] : 0; P.nlevels > 0 && k < P.tiny_ptr[1]; ++k) {
      const int s = P.tiny_idx[k];
      h->leaf_cap = std::max(h->leaf_cap, (P.col_ptr[s + 1] - P.col_ptr[s]) + (P.row_ptr[s + 1] - P.row_ptr[s]));
    }
  h->leaf_cap = std::min(SF_TBUF, (h->leaf_cap + 3) & ~3);
  h->arenaL.alloc(std::max<size_t>(totL, 1));
  h->arenaStack.alloc(std::max<size_t>(totS, 1));
  h->arenaBI.alloc(std::max<size_t>(totBI, 1));
  CK(cudaMemset(h->arenaBI.p, 0, std::max<size_t>(totBI, 1) * sizeof(int)));
  CK(cudaMemset(h->arenaStack.p, 0, std::max<size_t>(totS, 1) * sizeof(double)));
  std::vector<SparseBlock> sb(n_local);
  for (int f = 0; f < n_local; ++f) {
    const PatternPlan &P = h->plans[h->block_plan[f]];
    SparseBlock &B = sb[f];
    B.plan = h->block_plan[f];
    B.root = f;
    B.val_off = first_k[f] < 0 ? 0 : first_k[f];
    B.slot_base = nvals;
    B.L = h->arenaL.p + oL[f];
    B.cb = h->arenaStack.p + oS[f];
    B.vec = B.cb + P.cb_total + 8;
    int *bi = h->arenaBI.p + oBI[f];
    B.fid = bi; bi += P.fid_total;
    B.pbz = bi; bi += P.fs_total;
    B.opos = bi; bi += P.fs_total;
    B.meta = bi; bi += 3 * P.ns;
    B.rootids = bi; bi += P.nT + P.DR;
    B.info = bi;
  }
  h->blocks_dev.upload(sb);
  h->vec_off.upload(voff);
  h->root_off.upload(roff_root);
  {
    std::vector<int64_t> r64(roff_root.begin(), roff_root.end());
    h->root_off64.upload(r64);
  }
  h->ywork.alloc((size_t)std::max<int64_t>(h->local_dim, 1));
  h->root_rhs.alloc((size_t)std::max<int64_t>(h->root_total, 1));
  h->root_x.alloc((size_t)std::max<int64_t>(h->root_total, 1));

  // ---- coupling-row sources: row r <- (front, position) in front order ----
  std::vector<int32_t> brow((size_t)bptr[n_local]);
  std::vector<int64_t> sptr((size_t)m_c + 1, 0);
  for (int f = 0; f < n_local; ++f)
    for (int64_t p2 = border_ptr[f]; p2 < border_ptr[f + 1]; ++p2) {
      brow[(size_t)(bptr[f] + (p2 - border_ptr[f]))] = border_rows[p2];
      sptr[(size_t)border_rows[p2] + 1]++;
    }
  for (int r = 0; r < m_c; ++r) sptr[r + 1] += sptr[r];
  std::vector<int32_t> sfront((size_t)bptr[n_local]), spos((size_t)bptr[n_local]);
  std::vector<int64_t> fill(sptr.begin(), sptr.end() - 1);
  for (int f = 0; f < n_local; ++f)
    for (int a = 0; a < h->m[f]; ++a) {
      const int r = brow[(size_t)bptr[f] + a];
      sfront[(size_t)fill[r]] = f;
      spos[(size_t)fill[r]] = a;
      fill[r]++;
    }
  h->brow.upload(brow);
  h->brow_ptr.upload(bptr);
  h->src_ptr.upload(sptr);
  h->src_front.upload(sfront);
  h->src_pos.upload(spos);
  {
    std::vector<int64_t> aoff(spos.size());
    std::vector<int32_t> sld(spos.size());
    for (size_t p = 0; p < spos.size(); ++p) {
      const int f = sfront[p];
      const int nbf = h->n[(size_t)f];  // first border row of the front (static)
      aoff[p] = (int64_t)offA[(size_t)f] + nbf + spos[p] + (int64_t)nbf * h->ld[(size_t)f];
      sld[p] = h->m[(size_t)f] == m_c ? h->ld[(size_t)f] : -h->ld[(size_t)f];
    }
    h->src_aoff.upload(aoff);
    h->src_ld.upload(sld);
  }
  {
    std::vector<int64_t> boff(spos.size());
    for (size_t p = 0; p < spos.size(); ++p)
      boff[p] = (int64_t)offZ[(size_t)sfront[p]] + h->n[(size_t)sfront[p]] + spos[p];
    h->src_boff.upload(boff);
  }

  // ---- row lists of K for the residual (iterative refinement) ----
  {
    const int64_t ldim = h->local_dim;
    std::vector<std::vector<std::pair<int, long long>>> rl((size_t)ldim), rb((size_t)m_c), rq((size_t)m_c);
    std::vector<int64_t> boff(n_local + 1, 0);
    for (int f = 0; f < n_local; ++f) boff[f + 1] = boff[f] + block_n[f];
    for (int64_t k = 0; k < nvals; ++k) {
      const int f = dest_front[k];
      if (f < 0) continue;
      const int r = dest_row[k], c = dest_col[k];
      if (f == n_local) {
        rq[(size_t)r].push_back({c, k});
        if (r != c) rq[(size_t)c].push_back({r, k});
      } else if (r < block_n[f]) {
        const int64_t a = boff[f] + r, b = boff[f] + c;
        rl[(size_t)a].push_back({(int)b, k});
        if (a != b) rl[(size_t)b].push_back({(int)a, k});
      } else {
        const int g = border_rows[border_ptr[f] + (r - block_n[f])];
        const int64_t b = boff[f] + c;
        rl[(size_t)b].push_back({(int)(ldim + g), k});
        rb[(size_t)g].push_back({(int)b, k});
      }
    }
    if (h->have_classes) {
      for (int64_t a = 0; a < ldim; ++a)
        if (h->cls_local[(size_t)a] > 0) rl[(size_t)a].push_back({(int)a, nvals + (h->cls_local[(size_t)a] - 1)});
      for (int g = 0; g < m_c; ++g)
        if (h->cls_c[(size_t)g] > 0) rq[(size_t)g].push_back({g, nvals + (h->cls_c[(size_t)g] - 1)});
    }
    auto flat = [](const std::vector<std::vector<std::pair<int, long long>>> &rows, DevBuf<long long> &ptr,
                   DevBuf<int> &col, DevBuf<long long> &src) {
      std::vector<long long> hp(rows.size() + 1, 0), hs;
      std::vector<int> hc;
      for (size_t i = 0; i < rows.size(); ++i) {
        for (auto &e : rows[i]) { hc.push_back(e.first); hs.push_back(e.second); }
        hp[i + 1] = (long long)hc.size();
      }
      if (hc.empty()) { hc.push_back(0); hs.push_back(0); }
      ptr.upload(hp);
      col.upload(hc);
      src.upload(hs);
    };
    if (ldim + m_c > 0x7fffffffLL) return fail("pp_symbolic: local dimension exceeds 2^31");
    flat(rl, h->rl_ptr, h->rl_col, h->rl_src);
    flat(rb, h->rb_ptr, h->rb_col, h->rb_src);
    flat(rq, h->rq_ptr, h->rq_col, h->rq_src);
    h->res_blocks = (int)((ldim + 255) / 256);
    h->res_loc.alloc((size_t)std::max<int64_t>(ldim, 1));
    h->dx_tmp.alloc((size_t)std::max<int64_t>(ldim, 1));
    h->res_c.alloc((size_t)std::max(m_c, 1));
    h->dxc_tmp.alloc((size_t)std::max(m_c, 1));
    h->bc_keep.alloc((size_t)std::max(m_c, 1));
    h->res_part.alloc((size_t)2 * std::max(h->res_blocks, 1));
    h->res_out.alloc(2);
    h->solved = false;
    h->staged_from = nullptr;
  }

  // ---- solve buffers ----
  std::vector<int64_t> zero_off(nfronts + 1, 0);
  h->rhs_off.upload(zero_off);
  h->rhs.alloc((size_t)std::max<int64_t>(h->local_dim, 1));
  h->x.alloc((size_t)std::max<int64_t>(h->local_dim, 1));
  h->xc.alloc((size_t)std::max(m_c, 1));
  h->crhs.alloc((size_t)std::max(m_c, 1));
  h->bytes = (int64_t)((totA + totW + totZ + totL + totS) * sizeof(double) + (totI + totBI) * sizeof(int));
  h->schur_size = h->child ? h->sc_nnz : (int64_t)m_c * m_c;
  if (h->depth > 0) {  // a child keeps its own Schur and coupling-rhs buffers (nothing is reduced across ranks)
    h->own_schur.alloc((size_t)h->schur_size + PP_SCHUR_TAIL);
    h->own_rc.alloc((size_t)std::max(m_c, 1));
  }
  h->have_symbolic = true;
  return (int)PP_SUCCESSFUL;
}

// Decides how the coupling system is held -- one dense front, or (sparse pattern, e.g. the block-tridiagonal S of a
// chain of time blocks) a child handle that factorises S as a block-bordered matrix of its own -- and, in the second
// case, builds the pattern tables and runs the child's symbolic phase (which may recurse).
// Every rank reaches the same decision: it only depends on m_c, the borders of ALL blocks and the pattern of Q.
static int setup_coupling(pp_handle *h) {
  delete h->child;
  h->child = nullptr;
  h->sc_nnz = 0;
  const int m_c = h->m_c, n_local = h->n_local;
  if (m_c == 0 || h->depth >= h->cpl_opt.max_levels) return PP_SUCCESSFUL;
  std::vector<int32_t> qrow, qcol;
  std::vector<int64_t> qk;
  for (int64_t k = 0; k < h->nvals; ++k)
    if (h->in_dest_front[(size_t)k] == n_local) {
      const int r = h->in_dest_row[(size_t)k], c = h->in_dest_col[(size_t)k];
      if (r < 0 || c < 0 || r >= m_c || c > r) return misuse("pp_symbolic: coupling entry outside the lower triangle");
      qrow.push_back(r);
      qcol.push_back(c);
      qk.push_back(k);
    }
  const std::vector<int64_t> &cp = h->have_cliques ? h->clq_ptr : h->in_border_ptr;
  const std::vector<int32_t> &cr = h->have_cliques ? h->clq_rows : h->in_border_rows;
  for (int32_t r : cr)
    if (r < 0 || r >= m_c) return misuse("pp_symbolic: clique row out of range");
  CouplingLevel L = analyse_coupling(m_c, cp, cr, qrow, qcol, h->cpl_opt);
  if (!L.sparse) return PP_SUCCESSFUL;
  const int64_t nnz = L.nnz();
  // pattern entries on the device, Q sources per entry
  std::vector<int32_t> er((size_t)nnz), ec((size_t)nnz);
  for (int c = 0; c < m_c; ++c)
    for (int64_t p = L.colptr[(size_t)c]; p < L.colptr[(size_t)c + 1]; ++p) { er[(size_t)p] = L.rowidx[(size_t)p]; ec[(size_t)p] = c; }
  std::vector<std::pair<int64_t, int64_t>> qs;  // (pattern slot, value index), input order inside a slot
  qs.reserve(qk.size());
  for (size_t i = 0; i < qk.size(); ++i) {
    const int64_t a = L.colptr[(size_t)qcol[i]], b = L.colptr[(size_t)qcol[i] + 1];
    const auto it = std::lower_bound(L.rowidx.begin() + a, L.rowidx.begin() + b, qrow[i]);
    qs.emplace_back((int64_t)(it - L.rowidx.begin()), qk[i]);
  }
  if (h->have_classes && (int)h->cls_c.size() == m_c)
    for (int g = 0; g < m_c; ++g)
      if (h->cls_c[(size_t)g] > 0)   // the diagonal is the first entry of column g
        qs.emplace_back(L.colptr[(size_t)g], h->nvals + (h->cls_c[(size_t)g] - 1));
  std::stable_sort(qs.begin(), qs.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
  std::vector<int64_t> qptr((size_t)nnz + 1, 0), qsrc(qs.size());
  for (size_t i = 0; i < qs.size(); ++i) { qptr[(size_t)qs[i].first + 1]++; qsrc[i] = qs[i].second; }
  for (int64_t e = 0; e < nnz; ++e) qptr[(size_t)e + 1] += qptr[(size_t)e];
  if (qsrc.empty()) qsrc.push_back(0);
  h->sp_row.upload(er);
  h->sp_col.upload(ec);
  h->sp_qptr.upload(qptr);
  h->sp_qsrc.upload(qsrc);
  h->sp_perm_local.upload(L.perm_local);
  if (L.perm_c.empty()) L.perm_c.push_back(0);
  h->sp_perm_c.upload(L.perm_c);
  h->sp_vals.alloc((size_t)nnz);
  h->sc_nnz = nnz;
  // the child: S as a block-bordered matrix (dense diagonal blocks, replicated on every rank)
  pp_handle *c = nullptr;
  int rc = pp_create(h->device, &c);
  if (rc != PP_SUCCESSFUL) return rc;
  h->child = c;
  c->depth = h->depth + 1;
  c->cpl_opt = h->cpl_opt;
  c->pivot_tol = h->pivot_tol;
  c->pivot_threshold = h->pivot_threshold;
  c->panel_width = h->panel_width;
  c->use_cluster = h->use_cluster;
  c->panel_onchip = h->panel_onchip;
  c->panel_spec = h->panel_spec;
  c->use_small = h->use_small;
  c->overlap_groups = h->overlap_groups;
  c->update_strip = h->update_strip;
  c->use_sparse = false;     // the blocks of S are dense
  c->defer_status = 1;
  rc = pp_symbolic(c, L.n_blocks, L.block_n.data(), L.border_ptr.data(), L.border_rows.data(), L.m_next, nnz,
                   L.dest_front.data(), L.dest_row.data(), L.dest_col.data(), nullptr);
  if (rc == PP_MISUSE) return fail(std::string("internal error in the coupling analysis: ") + g_error);
  return rc;
}

int pp_symbolic(pp_handle *h, int32_t n_local, const int32_t *block_n, const int64_t *border_ptr,
                const int32_t *border_rows, int32_t m_c, int64_t nvals, const int32_t *dest_front,
                const int32_t *dest_row, const int32_t *dest_col, const double *values_hint) {
  if (!h) return misuse("pp_symbolic: null handle");
  if (n_local < 0 || m_c < 0 || nvals < 0) return misuse("pp_symbolic: negative size");
  if (n_local > 0 && (!block_n || !border_ptr)) return misuse("pp_symbolic: null block description");
  if (nvals > 0 && (!dest_front || !dest_row || !dest_col)) return misuse("pp_symbolic: null destination arrays");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    h->have_symbolic = false;
    for (int f = 0; f < n_local; ++f) {
      if (block_n[f] < 0) return misuse("pp_symbolic: negative block order");
      const int64_t mi = border_ptr[f + 1] - border_ptr[f];
      if (mi < 0 || mi > m_c) return misuse("pp_symbolic: bad border_ptr");
      for (int64_t p = border_ptr[f]; p < border_ptr[f + 1]; ++p) {
        if (border_rows[p] < 0 || border_rows[p] >= m_c) return misuse("pp_symbolic: border row out of range");
        if (p > border_ptr[f] && border_rows[p] <= border_rows[p - 1])
          return misuse("pp_symbolic: border rows must be strictly ascending");
      }
    }
    h->n_local = n_local;
    h->m_c = m_c;
    h->nvals = nvals;
    h->in_block_n.assign(block_n, block_n + n_local);
    h->in_border_ptr.assign(border_ptr, border_ptr + n_local + 1);
    if (n_local == 0) h->in_border_ptr.assign(1, 0);
    h->in_border_rows.assign(border_rows, border_rows + (n_local ? border_ptr[n_local] : 0));
    h->in_dest_front.assign(dest_front, dest_front + nvals);
    h->in_dest_row.assign(dest_row, dest_row + nvals);
    h->in_dest_col.assign(dest_col, dest_col + nvals);
    if (values_hint) h->in_hint.assign(values_hint, values_hint + nvals); else h->in_hint.clear();
    h->sparse_failed = false;
    const int rc = setup_coupling(h);
    if (rc != PP_SUCCESSFUL) return rc;
    return do_symbolic(h, false);
  });
}

int pp_set_coupling_cliques(pp_handle *h, int32_t n_cliques, const int64_t *ptr, const int32_t *rows) {
  if (!h) return misuse("pp_set_coupling_cliques: null handle");
  if (n_cliques < 0 || (n_cliques > 0 && (!ptr || (ptr[n_cliques] > 0 && !rows))))
    return misuse("pp_set_coupling_cliques: null argument");
  return guarded([&]() {
    h->have_cliques = n_cliques > 0;
    h->clq_ptr.clear();
    h->clq_rows.clear();
    if (n_cliques > 0) {
      for (int k = 0; k < n_cliques; ++k) {
        if (ptr[k + 1] < ptr[k]) return misuse("pp_set_coupling_cliques: bad ptr");
        for (int64_t p = ptr[k] + 1; p < ptr[k + 1]; ++p)
          if (rows[p] <= rows[p - 1]) return misuse("pp_set_coupling_cliques: rows must be strictly ascending");
      }
      h->clq_ptr.assign(ptr, ptr + n_cliques + 1);
      h->clq_rows.assign(rows, rows + ptr[n_cliques]);
    }
    return (int)PP_SUCCESSFUL;
  });
}

int64_t pp_schur_size(const pp_handle *h) { return h ? h->schur_size : 0; }

int pp_set_diagonal_classes(pp_handle *h, int64_t n_local_rows, const int8_t *cls_local, int32_t m_c,
                            const int8_t *cls_c) {
  if (!h) return misuse("pp_set_diagonal_classes: null handle");
  if (n_local_rows < 0 || m_c < 0 || (n_local_rows > 0 && !cls_local && cls_c) || (m_c > 0 && !cls_c && cls_local))
    return misuse("pp_set_diagonal_classes: bad argument");
  return guarded([&]() {
    h->have_classes = cls_local != nullptr || cls_c != nullptr;
    h->cls_local.clear();
    h->cls_c.clear();
    if (h->have_classes) {
      for (int64_t i = 0; i < n_local_rows; ++i)
        if (cls_local[i] < 0 || cls_local[i] > PP_SHIFT_SLOTS) return misuse("pp_set_diagonal_classes: class out of range");
      for (int i = 0; i < m_c; ++i)
        if (cls_c[i] < 0 || cls_c[i] > PP_SHIFT_SLOTS) return misuse("pp_set_diagonal_classes: class out of range");
      h->cls_local.assign(cls_local, cls_local + n_local_rows);
      h->cls_c.assign(cls_c, cls_c + m_c);
    }
    h->have_symbolic = false;  // the plans must be rebuilt with the classified diagonal entries
    for (int k = 0; k < PP_SHIFT_SLOTS; ++k) h->shifts[k] = 0.0;
    return (int)PP_SUCCESSFUL;
  });
}

int pp_set_shifts(pp_handle *h, const double *shifts) {
  if (!h || !shifts) return misuse("pp_set_shifts: null argument");
  for (int k = 0; k < PP_SHIFT_SLOTS; ++k) {
    if (!std::isfinite(shifts[k])) return misuse("pp_set_shifts: non-finite shift");
    if (shifts[k] != 0.0 && !h->have_classes) return misuse("pp_set_shifts: pp_set_diagonal_classes required first");
    h->shifts[k] = shifts[k];
  }
  return PP_SUCCESSFUL;
}

int64_t pp_value_uploads(const pp_handle *h) { return h ? h->value_uploads : 0; }

int pp_coupling_stats(pp_handle *h, int64_t out[8]) {
  if (!h || !h->have_symbolic || !out) return misuse("pp_coupling_stats: bad argument");
  int levels = 0;
  int64_t last_mc = h->m_c, blocks = 0, maxfront = 0;
  for (pp_handle *c = h->child; c; c = c->child) {
    ++levels;
    last_mc = c->m_c;
    blocks += c->n_local;
    maxfront = std::max<int64_t>(maxfront, c->nfmax_local);
  }
  out[0] = levels;          // child levels below this handle (0 = dense coupling front)
  out[1] = h->schur_size;   // doubles of the Schur payload the caller all-reduces
  out[2] = h->m_c;
  out[3] = last_mc;         // order of the dense coupling front at the bottom of the chain
  out[4] = blocks;          // diagonal blocks over all child levels
  out[5] = maxfront;        // largest front of the child levels
  out[6] = h->child ? h->child->n_local : 0;
  out[7] = 0;
  return PP_SUCCESSFUL;
}

static int numeric_local_once(pp_handle *h, const double *dvals, double *schur_local_dev, cudaStream_t st,
                              int *sparse_bad, bool defer = false) {
  {
    ProfSpan sp(h, PP_PROF_ASSEMBLE, st);
    CK(cudaMemsetAsync(h->arenaA.p, 0, h->arenaA_elems * sizeof(double), st));
    reset_fronts_kernel<<<h->n_local + 1, 256, 0, st>>>(h->fronts.p, h->inertia.p);
    h->launches++;
    if (h->nuniq > 0) {
      assemble_kernel<<<(unsigned)((h->nuniq + 255) / 256), 256, 0, st>>>(dvals, h->asm_dst.p, h->asm_ptr.p,
                                                                        h->asm_src.p, h->nuniq, h->arenaA.p);
      h->launches++;
    }
  }
  if (h->n_local > 0) {
    ProfSpan sp(h, PP_PROF_SUBTREE, st);
    if (h->max_leaves > 0) {
      const int lg = h->leaf_cap <= 8 ? 8 : (h->leaf_cap <= 16 ? 16 : 32), per = LF_NT / lg;
      dim3 g((h->max_leaves + per - 1) / per, h->n_local);
      const size_t lsm = per * align16(fb_bytes(h->leaf_cap, h->leaf_cap | 1)) + 16;
      auto kern = lg == 8 ? subtree_leaf_kernel<8> : (lg == 16 ? subtree_leaf_kernel<16> : subtree_leaf_kernel<32>);
      kern<<<g, LF_NT, lsm, st>>>(h->blocks_dev.p, h->plans_dev.p, dvals, h->pivot_threshold, h->pivot_tol,
                                  h->inertia.p, h->leaf_cap);
      h->launches++;
    }
    launch_clustered(subtree_factor_kernel, h->n_local, subtree_csize(h, true), SF_NT, SF_SMEM, st, h->blocks_dev.p,
                     h->plans_dev.p, h->fronts.p, dvals, h->pivot_threshold, h->pivot_tol, h->inertia.p);
    h->launches++;
  }
  factor_fronts(h, 0, h->n_local, st);
  if (h->m_c > 0 && h->child) {
    ProfSpan sp(h, PP_PROF_SCHUR, st);
    schur_gather_sparse_kernel<<<(unsigned)((h->sc_nnz + 255) / 256), 256, 0, st>>>(
        h->arenaA.p, h->src_ptr.p, h->src_front.p, h->src_pos.p, h->src_aoff.p, h->src_ld.p, h->brow_ptr.p, h->brow.p,
        h->sp_row.p, h->sp_col.p, h->sc_nnz, schur_local_dev);
    h->launches++;
  } else if (h->m_c > 0) {
    ProfSpan sp(h, PP_PROF_SCHUR, st);
    const bool warp_per_entry = h->m_c <= 256;   // few entries with many sources each: a warp per entry
    dim3 g(warp_per_entry ? (h->m_c + 7) / 8 : (h->m_c + 127) / 128, h->m_c);
    (warp_per_entry ? schur_gather_warp_kernel : schur_gather_kernel)<<<g, warp_per_entry ? 256 : 128, 0, st>>>(h->fronts.p, h->arenaA.p, h->src_ptr.p, h->src_front.p, h->src_pos.p,
                                           h->src_aoff.p, h->src_ld.p, h->brow_ptr.p, h->brow.p, h->m_c,
                                           schur_local_dev);
    h->launches++;
  }
  double *tail = schur_local_dev ? schur_local_dev + (size_t)h->schur_size : nullptr;
  int bad = 0;
  *sparse_bad = 0;
  if (h->n_local > 0) {
    // inertia of the roots, status flags and the Schur tail in one launch (the last CTA to finish packs them)
    local_finalize_kernel<<<h->n_local, 256, 0, st>>>(h->fronts.p, h->blocks_dev.p, h->n_local, h->inertia.p, h->flag.p,
                                                      tail);
    h->launches++;
    CK(cudaGetLastError());
    if (!defer) {
      h->pin_flag.ensure(8);
      CK(cudaMemcpyAsync(h->pin_flag.p, h->flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      bad = h->pin_flag.p[0];
      *sparse_bad = h->pin_flag.p[1];
    }
  } else {
    CK(cudaMemsetAsync(h->flag.p, 0, 2 * sizeof(int), st));
    if (tail) {
      pack_tail_kernel<<<1, 32, 0, st>>>(tail, h->flag.p, h->inertia.p);
      h->launches++;
    }
  }
  return bad;
}

int pp_numeric_local(pp_handle *h, const double *values, int on_device, double *schur_local_dev,
                     void *stream) {
  if (!h || !h->have_symbolic) return misuse("pp_numeric_local: symbolic factorization required first");
  if (h->nvals > 0 && !values && on_device != PP_VALUES_REUSE) return misuse("pp_numeric_local: null values");
  if (!schur_local_dev) return misuse("pp_numeric_local: null schur buffer");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    h->local_factored = h->coupling_factored = h->forward_done = false;
    const double *dvals = values;
    if (on_device == PP_VALUES_REUSE) {
      if (!h->last_vals) return misuse("pp_numeric_local: no previous values to reuse");
      dvals = h->last_vals;   // a retry of the inertia-correction loop: same values, new shifts
    } else if (!on_device && h->nvals > 0 && h->staged_from == (const void *)values) {
      dvals = h->vals.p;  // pp_stage_values gathered these values and already queued their transfer
      h->value_uploads++;
    } else if (!on_device && h->nvals > 0) {
      const double *src = values;
      if (!is_pinned_host(values)) {  // pageable caller memory: stage through the handle's pinned buffer
        h->pin_vals.ensure((size_t)h->nvals);
        std::memcpy(h->pin_vals.p, values, (size_t)h->nvals * sizeof(double));
        src = h->pin_vals.p;
      }
      CK(cudaMemcpyAsync(h->vals.p, src, (size_t)h->nvals * sizeof(double), cudaMemcpyHostToDevice, st));
      dvals = h->vals.p;
      h->value_uploads++;
    } else if (on_device && h->have_classes && h->nvals > 0 && values != h->vals.p) {
      // the shift slots live behind the handle's own copy of the values
      CK(cudaMemcpyAsync(h->vals.p, values, (size_t)h->nvals * sizeof(double), cudaMemcpyDeviceToDevice, st));
      dvals = h->vals.p;
    }
    if (h->have_classes) {
      if (dvals != h->vals.p) return misuse("pp_numeric_local: diagonal shifts need the handle's copy of the values");
      h->pin_shifts.ensure(PP_SHIFT_SLOTS);
      for (int k = 0; k < PP_SHIFT_SLOTS; ++k) h->pin_shifts.p[k] = h->shifts[k];
      CK(cudaMemcpyAsync(h->vals.p + h->nvals, h->pin_shifts.p, PP_SHIFT_SLOTS * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    h->staged_from = nullptr;
    h->tail_valid = false;
    h->last_vals = dvals;
    h->solved = false;
    h->inertia_cached = false;
    h->status_pending = false;
    h->last_schur = schur_local_dev;
    int sparse_bad = 0;
    if (h->defer_status) {
      numeric_local_once(h, dvals, schur_local_dev, st, &sparse_bad, true);
      h->status_pending = h->defer_status == 1;
      h->local_factored = true;
      return (int)PP_SUCCESSFUL;  // provisional: pp_numeric_coupling reports the final status
    }
    int bad = numeric_local_once(h, dvals, schur_local_dev, st, &sparse_bad);
    if (sparse_bad && h->no_fallback) return fail("pp_numeric_local: sparse path overflow (fallback disabled)");
    if (sparse_bad) {
      // A block ran out of delayed-pivot capacity (or front buffer): redo the analysis with whole
      // blocks as dense fronts -- slower, but pivoting is then unrestricted -- and factor again.
      h->sparse_failed = true;
      const int rc = do_symbolic(h, true);
      if (rc != PP_SUCCESSFUL) return rc;
      bad = numeric_local_once(h, dvals, schur_local_dev, st, &sparse_bad);
      if (sparse_bad) return fail("pp_numeric_local: internal error in the dense re-factorisation");
    }
    h->local_factored = true;
    return bad ? (int)PP_SINGULAR : (int)PP_SUCCESSFUL;
  });
}

// ---- coupling phase -----------------------------------------------------------------------------------
static void enqueue_numeric_all(pp_handle *c, const double *dvals, cudaStream_t st);

// S = Q + reduced Schur sum, factorised: a dense front, or (sparse pattern) the child handle, level by level.
// Nothing is synchronised; the statuses are read afterwards (chain_status).
static void enqueue_coupling(pp_handle *h, const double *schur_sum_dev, cudaStream_t st) {
  const int mc = h->m_c;
  if (mc == 0) return;
  if (h->child) {
    coupling_values_kernel<<<(unsigned)((h->sc_nnz + 255) / 256), 256, 0, st>>>(
        schur_sum_dev, h->last_vals, h->sp_qptr.p, h->sp_qsrc.p, h->sc_nnz, h->sp_vals.p);
    h->launches++;
    CK(cudaGetLastError());
    enqueue_numeric_all(h->child, h->sp_vals.p, st);
    return;
  }
  const Front C = h->hfronts[h->n_local];
  if (h->use_small && mc <= SM_CAP) {
    // small coupling system: S = Q + sum, its LDL^T and its inertia in one launch
    ProfSpan sp(h, PP_PROF_PANEL, st);
    front_small_kernel<<<1, SF_NT, SM_SMEM, st>>>(h->fronts.p + h->n_local, h->pivot_threshold, h->pivot_tol,
                                                  schur_sum_dev, mc, h->inertia.p + 3);
    h->launches++;
    CK(cudaGetLastError());
    return;
  }
  dim3 g((mc + 127) / 128, mc);
  coupling_add_kernel<<<g, 128, 0, st>>>(C, schur_sum_dev, mc);
  h->launches++;
  factor_fronts(h, h->n_local, 1, st);
  CK(cudaMemsetAsync(h->inertia.p + 3, 0, 3 * sizeof(unsigned long long), st));
  front_inertia_kernel<<<1, 256, 0, st>>>(h->fronts.p + h->n_local, h->inertia.p + 3);
  h->launches++;
  CK(cudaGetLastError());
}

// A child level: local blocks, its own Schur complement, its coupling system; flags and inertia counters are sent
// to pinned memory on the same stream (the parent's one synchronisation completes them).
static void enqueue_numeric_all(pp_handle *c, const double *dvals, cudaStream_t st) {
  int sparse_bad = 0;
  c->last_vals = dvals;
  c->solved = false;
  c->local_factored = c->coupling_factored = c->forward_done = false;
  numeric_local_once(c, dvals, c->own_schur.p, st, &sparse_bad, true);
  c->local_factored = true;
  enqueue_coupling(c, c->own_schur.p, st);
  post_flag(c, c->n_local, (c->m_c > 0 && !c->child) ? 1 : 0, 4, st);
  c->pin_flag.ensure(8);
  c->pin_inertia.ensure(8);
  CK(cudaMemcpyAsync(c->pin_flag.p, c->flag.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(c->pin_inertia.p, c->inertia.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  c->coupling_factored = true;
}

// After the stream was synchronised: status and inertia of the whole chain below h (the factorisation of S).
static void chain_status(pp_handle *h) {
  h->chain_singular = false;
  for (int k = 0; k < 3; ++k) h->chain_inertia[k] = 0;
  for (pp_handle *c = h->child; c; c = c->child) {
    if (c->pin_flag.p[0] || c->pin_flag.p[1] || c->pin_flag.p[4]) h->chain_singular = true;
    for (int k = 0; k < 3; ++k) h->chain_inertia[k] += c->pin_inertia.p[k];
    if (!c->child)
      for (int k = 0; k < 3; ++k) h->chain_inertia[k] += c->pin_inertia.p[3 + k];
  }
}

int pp_numeric_coupling(pp_handle *h, const double *schur_sum_dev, void *stream) {
  if (!h || !h->local_factored) return misuse("pp_numeric_coupling: pp_numeric_local required first");
  if (h->m_c > 0 && !schur_sum_dev) return misuse("pp_numeric_coupling: null schur buffer");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    h->coupling_factored = h->forward_done = false;
    const int mc = h->m_c;
    const int dense_c = (mc > 0 && !h->child) ? 1 : 0;  // a dense coupling front reports through its own info word
    enqueue_coupling(h, schur_sum_dev, st);
    if (h->status_pending) {
      // single-rank fast path: the local status, the coupling status and both inertias in ONE synchronisation
      if (schur_sum_dev != h->last_schur) return misuse("pp_numeric_coupling: defer_status needs the local Schur buffer");
      h->status_pending = false;
      post_flag(h, h->n_local, dense_c, 4, st);
      fetch_status(h, st);
      if (h->n_local > 0 && h->pin_flag.p[1]) {  // sparse path overflow: redo with dense blocks, synchronously
        if (h->no_fallback) return fail("pp_numeric_local: sparse path overflow (fallback disabled)");
        h->sparse_failed = true;
        const int rc = do_symbolic(h, true);
        if (rc != PP_SUCCESSFUL) return rc;
        int sparse_bad = 0;
        const int bad_local = numeric_local_once(h, h->last_vals, h->last_schur, st, &sparse_bad);
        if (sparse_bad) return fail("pp_numeric_local: internal error in the dense re-factorisation");
        h->local_factored = true;
        if (bad_local) return (int)PP_SINGULAR;
        enqueue_coupling(h, schur_sum_dev, st);
        post_flag(h, h->n_local, dense_c, 4, st);
        fetch_status(h, st);
      }
      chain_status(h);
      for (int k = 0; k < 6; ++k) h->inertia_cache[k] = h->pin_inertia.p[k];
      h->inertia_cached = true;
      if (h->pin_flag.p[0]) return (int)PP_SINGULAR;
      h->coupling_factored = true;
      return (h->pin_flag.p[4] || h->chain_singular) ? (int)PP_SINGULAR : (int)PP_SUCCESSFUL;
    }
    if (h->defer_status == 2 && schur_sum_dev) {
      // several ranks: the reduced tail of the Schur buffer (status of every rank's local phase, summed inertia),
      // this rank's coupling status and its inertia counters come back with ONE synchronisation
      post_flag(h, h->n_local, dense_c, 4, st);
      h->pin_tail.ensure(PP_SCHUR_TAIL);
      CK(cudaMemcpyAsync(h->pin_tail.p, schur_sum_dev + (size_t)h->schur_size, PP_SCHUR_TAIL * sizeof(double),
                         cudaMemcpyDeviceToHost, st));
      fetch_status(h, st);
      chain_status(h);
      h->tail_valid = true;
      for (int k = 0; k < 6; ++k) h->inertia_cache[k] = h->pin_inertia.p[k];
      h->inertia_cached = true;
      for (int k = 0; k < PP_SCHUR_TAIL; ++k)
        if (!std::isfinite(h->pin_tail.p[k])) return fail("pp_numeric_coupling: non-finite status tail in the Schur buffer");
      if (h->pin_tail.p[1] > 0.0) return (int)PP_NOT_ENOUGH_MEMORY;  // some rank's sparse path overflowed: redo densely
      if (h->pin_tail.p[0] > 0.0) return (int)PP_SINGULAR;
      h->coupling_factored = true;
      return (h->pin_flag.p[4] || h->chain_singular) ? (int)PP_SINGULAR : (int)PP_SUCCESSFUL;
    }
    const int bad = read_flag(h, h->n_local, dense_c, st) | 0;
    if (!dense_c) CK(cudaStreamSynchronize(st));
    chain_status(h);
    h->coupling_factored = true;
    return (bad || h->chain_singular) ? (int)PP_SINGULAR : (int)PP_SUCCESSFUL;
  });
}

int pp_schur_tail(pp_handle *h, double out[8]) {
  if (!h || !out) return misuse("pp_schur_tail: null argument");
  if (!h->tail_valid) return misuse("pp_schur_tail: available after pp_numeric_coupling with defer_status = 2");
  for (int k = 0; k < PP_SCHUR_TAIL; ++k) out[k] = h->pin_tail.p[k];
  return PP_SUCCESSFUL;
}

static int read_inertia(pp_handle *h, int which, int64_t out[3]) {
  return guarded([&]() {
    if (which == 1 && h->child) {  // sparse coupling system: the sum over the chain of child levels
      for (int k = 0; k < 3; ++k) out[k] = (int64_t)h->chain_inertia[k];
      return (int)PP_SUCCESSFUL;
    }
    if (h->inertia_cached) {
      for (int k = 0; k < 3; ++k) out[k] = (int64_t)h->inertia_cache[which * 3 + k];
      return (int)PP_SUCCESSFUL;
    }
    DeviceGuard dev_guard(h->device);
    h->pin_inertia.ensure(8);
    CK(cudaMemcpy(h->pin_inertia.p, h->inertia.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) out[k] = (int64_t)h->pin_inertia.p[which * 3 + k];
    return (int)PP_SUCCESSFUL;
  });
}

int pp_inertia_local(pp_handle *h, int64_t out[3]) {
  if (!h || !out) return misuse("pp_inertia_local: null argument");
  if (!h->local_factored) return misuse("pp_inertia_local: numeric factorization required first");
  return read_inertia(h, 0, out);
}

int pp_inertia_coupling(pp_handle *h, int64_t out[3]) {
  if (!h || !out) return misuse("pp_inertia_coupling: null argument");
  if (!h->coupling_factored) return misuse("pp_inertia_coupling: numeric factorization required first");
  return read_inertia(h, 1, out);
}

static void enqueue_residual_local(pp_handle *h, double *buf_dev, cudaStream_t st);
static void enqueue_residual_norms(pp_handle *h, const double *buf_sum_dev, cudaStream_t st);

// CTAs per front for the dense triangular sweeps: one CTA when there are enough fronts to fill the GPU (or the
// fronts are short), a cluster of up to 8 when a few tall fronts would otherwise stream their factors through a few
// SMs; at least 256 rows per CTA, and as many CTAs as the solve vector needs to fit in shared memory.
static int solve_csize(const pp_handle *h, int count, int nfmax) {
  int c = 1;
  if (nfmax >= 1024)
    while (c < CS_MAXC && count * c * 2 <= h->sm_count && nfmax / (c * 2) >= 256) c *= 2;
  while (c < CS_MAXC && solve_smem(nfmax) > h_optin_smem && cluster_solve_smem(nfmax, c) > h_optin_smem) c *= 2;
  return c;
}

static void run_forward(pp_handle *h, const double *drhs, double *rc_local_dev, cudaStream_t st) {
  if (h->n_local > 0) {
    ProfSpan sp(h, PP_PROF_FORWARD, st);
    if (h->max_leaves > 0) {
      const int lg = h->leaf_cap <= 8 ? 8 : (h->leaf_cap <= 16 ? 16 : 32), per = LF_NT / lg;
      dim3 g((h->max_leaves + per - 1) / per, h->n_local);
      const size_t lsm = per * align16(sb_bytes(h->leaf_cap, h->leaf_cap | 1)) + 16;
      auto kern = lg == 8 ? subtree_leaf_forward_kernel<8>
                          : (lg == 16 ? subtree_leaf_forward_kernel<16> : subtree_leaf_forward_kernel<32>);
      kern<<<g, LF_NT, lsm, st>>>(h->blocks_dev.p, h->plans_dev.p, drhs, h->vec_off.p, h->ywork.p, h->leaf_cap);
      h->launches++;
    }
    launch_clustered(subtree_forward_kernel, h->n_local, subtree_csize(h, false), SF_NT, SV_SMEM, st, h->blocks_dev.p,
                     h->plans_dev.p, drhs, h->vec_off.p, h->ywork.p, h->root_rhs.p, h->root_off.p);
    const int cs = solve_csize(h, h->n_local, h->nfmax_local);
    if (cs > 1)
      launch_clustered(front_forward_cluster_kernel, h->n_local, cs, CS_NT, cluster_solve_smem(h->nfmax_local, cs), st,
                       h->fronts.p, h->root_rhs.p, h->root_off64.p);
    else
      front_forward_kernel<512><<<h->n_local, 512, solve_smem(h->nfmax_local), st>>>(h->fronts.p, h->root_rhs.p,
                                                                                    h->root_off64.p);
    h->launches += 2;
  }
  if (h->m_c > 0) {
    rc_gather_kernel<<<(h->m_c + 3) / 4, 128, 0, st>>>(h->arenaZ.p, h->src_ptr.p, h->src_boff.p, h->m_c,
                                                          rc_local_dev);
    h->launches++;
  }
  CK(cudaGetLastError());
}

// x_c = S^-1 (drc + rc_sum) into dxc, then the backward sweeps into dx (all device pointers)
static void run_backward(pp_handle *h, const double *rc_sum_dev, const double *drc, double *dx, double *dxc,
                         cudaStream_t st) {
  const int mc = h->m_c;
  if (mc > 0 && h->child) {
    // sparse coupling system: S x_c = r_c + rc_sum is solved by the child level (which recurses the same way);
    // its local vector and its coupling vector are a permutation of this level's coupling variables
    pp_handle *c = h->child;
    const int nl = (int)c->local_dim, nc = c->m_c;
    if (nl > 0) gather_perm_kernel<<<(nl + 255) / 256, 256, 0, st>>>(drc, rc_sum_dev, h->sp_perm_local.p, nl, c->rhs.p);
    if (nc > 0) gather_perm_kernel<<<(nc + 255) / 256, 256, 0, st>>>(drc, rc_sum_dev, h->sp_perm_c.p, nc, c->crhs.p);
    h->launches += 2;
    run_forward(c, c->rhs.p, c->own_rc.p, st);
    run_backward(c, c->own_rc.p, c->crhs.p, c->x.p, c->xc.p, st);
    if (nl > 0) scatter_perm_kernel<<<(nl + 255) / 256, 256, 0, st>>>(c->x.p, h->sp_perm_local.p, nl, dxc);
    if (nc > 0) scatter_perm_kernel<<<(nc + 255) / 256, 256, 0, st>>>(c->xc.p, h->sp_perm_c.p, nc, dxc);
    h->launches += 2;
  } else if (mc > 0) {
    const int cs = solve_csize(h, 1, mc);
    if (cs > 1)
      launch_clustered(coupling_solve_cluster_kernel, 1, cs, CS_NT, cluster_solve_smem(mc, cs), st,
                       h->fronts.p + h->n_local, drc, rc_sum_dev, dxc);
    else
      coupling_solve_kernel<512><<<1, 512, solve_smem(mc), st>>>(h->fronts.p + h->n_local, drc, rc_sum_dev, dxc);
    h->launches++;
  }
  if (h->n_local > 0) {
    ProfSpan sp(h, PP_PROF_BACKWARD, st);
    const int cs = solve_csize(h, h->n_local, h->nfmax_local);
    if (cs > 1)
      launch_clustered(front_backward_cluster_kernel, h->n_local, cs, CS_NT, cluster_solve_smem(h->nfmax_local, cs), st,
                       h->fronts.p, dxc, h->brow_ptr.p, h->brow.p, h->root_x.p, h->root_off64.p);
    else
      front_backward_kernel<512><<<h->n_local, 512, solve_smem(h->nfmax_local), st>>>(
          h->fronts.p, dxc, h->brow_ptr.p, h->brow.p, h->root_x.p, h->root_off64.p);
    launch_clustered(subtree_backward_kernel, h->n_local, subtree_csize(h, false), SF_NT, SV_SMEM, st, h->blocks_dev.p,
                     h->plans_dev.p, h->ywork.p, h->vec_off.p, h->root_x.p, h->root_off.p, dx);
    if (h->max_leaves > 0) {
      const int lg = h->leaf_cap <= 8 ? 8 : (h->leaf_cap <= 16 ? 16 : 32), per = LF_NT / lg;
      dim3 g((h->max_leaves + per - 1) / per, h->n_local);
      const size_t lsm = per * align16(sb_bytes(h->leaf_cap, h->leaf_cap | 1)) + 16;
      auto kern = lg == 8 ? subtree_leaf_backward_kernel<8>
                          : (lg == 16 ? subtree_leaf_backward_kernel<16> : subtree_leaf_backward_kernel<32>);
      kern<<<g, LF_NT, lsm, st>>>(h->blocks_dev.p, h->plans_dev.p, h->ywork.p, h->vec_off.p, dx, h->leaf_cap);
      h->launches++;
    }
    h->launches += 2;
  }
  CK(cudaGetLastError());
}

// A second stream of the handle and the events to fork it from / join it into the caller's stream.
static cudaStream_t aux_stream(pp_handle *h, int g) {
  if (!h->ev_fork) CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  while ((int)h->aux_streams.size() <= g) {
    cudaStream_t s2;
    cudaEvent_t e2;
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    h->aux_streams.push_back(s2);
    h->aux_events.push_back(e2);
  }
  return h->aux_streams[g];
}

// `join`: work on another stream (the residual check) that the synchronisation has to cover as well
static void copy_out(pp_handle *h, const double *dx, const double *dxc, double *x_local, double *x_c,
                     cudaStream_t st, cudaEvent_t join = nullptr) {
  const int mc = h->m_c;
  // pinned caller buffers receive the device data directly; pageable ones go through the handle's staging buffer
  const bool direct = (h->local_dim == 0 || is_pinned_host(x_local)) && (mc == 0 || is_pinned_host(x_c));
  if (direct) {
    if (h->local_dim > 0)
      CK(cudaMemcpyAsync(x_local, dx, (size_t)h->local_dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (mc > 0) CK(cudaMemcpyAsync(x_c, dxc, (size_t)mc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (join) CK(cudaStreamWaitEvent(st, join, 0));
    CK(cudaStreamSynchronize(st));
    return;
  }
  h->pin_vec.ensure((size_t)h->local_dim + (size_t)mc);
  if (h->local_dim > 0)
    CK(cudaMemcpyAsync(h->pin_vec.p, dx, (size_t)h->local_dim * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (mc > 0)
    CK(cudaMemcpyAsync(h->pin_vec.p + h->local_dim, dxc, (size_t)mc * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (join) CK(cudaStreamWaitEvent(st, join, 0));
  CK(cudaStreamSynchronize(st));
  if (h->local_dim > 0) std::memcpy(x_local, h->pin_vec.p, (size_t)h->local_dim * sizeof(double));
  if (mc > 0) std::memcpy(x_c, h->pin_vec.p + h->local_dim, (size_t)mc * sizeof(double));
}

int pp_solve_forward(pp_handle *h, const double *rhs_local, int on_device, double *rc_local_dev,
                     void *stream) {
  if (!h || !h->local_factored) return misuse("pp_solve_forward: numeric factorization required first");
  if (h->local_dim > 0 && !rhs_local) return misuse("pp_solve_forward: null rhs");
  if (h->m_c > 0 && !rc_local_dev) return misuse("pp_solve_forward: null coupling buffer");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    h->solved = false;
    const double *drhs = rhs_local;
    if (!on_device && h->local_dim > 0) {
      const double *src = rhs_local;
      if (!is_pinned_host(rhs_local)) {  // pageable caller memory: stage through the handle's pinned buffer
        h->pin_vec.ensure((size_t)h->local_dim + (size_t)h->m_c);
        std::memcpy(h->pin_vec.p, rhs_local, (size_t)h->local_dim * sizeof(double));
        src = h->pin_vec.p;
      }
      CK(cudaMemcpyAsync(h->rhs.p, src, (size_t)h->local_dim * sizeof(double), cudaMemcpyHostToDevice, st));
      drhs = h->rhs.p;
    }
    h->last_rhs = drhs;
    run_forward(h, drhs, rc_local_dev, st);
    h->forward_done = true;
    return (int)PP_SUCCESSFUL;
  });
}

int pp_solve_backward(pp_handle *h, const double *rc_sum_dev, const double *rhs_c, int on_device,
                      double *x_local, double *x_c, void *stream) {
  if (!h || !h->coupling_factored || !h->forward_done)
    return misuse("pp_solve_backward: numeric factorization and pp_solve_forward required first");
  if (h->m_c > 0 && (!rc_sum_dev || !rhs_c || !x_c)) return misuse("pp_solve_backward: null coupling argument");
  if (h->local_dim > 0 && !x_local) return misuse("pp_solve_backward: null solution buffer");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int mc = h->m_c;
    double *dx = on_device ? x_local : h->x.p;
    double *dxc = on_device ? x_c : h->xc.p;
    if (mc > 0) {  // keep b_c on the device for the residual
      if (on_device) {
        CK(cudaMemcpyAsync(h->bc_keep.p, rhs_c, (size_t)mc * sizeof(double), cudaMemcpyDeviceToDevice, st));
      } else {
        h->pin_vec.ensure((size_t)h->local_dim + (size_t)mc);
        std::memcpy(h->pin_vec.p + h->local_dim, rhs_c, (size_t)mc * sizeof(double));
        CK(cudaMemcpyAsync(h->bc_keep.p, h->pin_vec.p + h->local_dim, (size_t)mc * sizeof(double), cudaMemcpyHostToDevice, st));
      }
    }
    run_backward(h, rc_sum_dev, h->bc_keep.p, dx, dxc, st);
    h->last_x = dx;
    h->last_xc = dxc;
    h->solved = true;
    h->norms_valid = false;
    cudaEvent_t join = nullptr;
    if (!on_device && h->auto_residual) {
      // single rank: nothing to reduce, so the residual norms ride on the synchronisation of the copy-out.  The check
      // only reads x: it runs on a second stream beside the device-to-host copy of x instead of in front of it.
      if (h->res_buf.n < (size_t)mc + 2) h->res_buf.alloc((size_t)mc + 2);
      cudaStream_t s2 = aux_stream(h, 0);
      CK(cudaEventRecord(h->ev_fork, st));
      CK(cudaStreamWaitEvent(s2, h->ev_fork, 0));
      enqueue_residual_local(h, h->res_buf.p, s2);
      enqueue_residual_norms(h, h->res_buf.p, s2);
      CK(cudaGetLastError());
      join = h->aux_events[0];
      CK(cudaEventRecord(join, s2));
    }
    if (!on_device) {
      copy_out(h, dx, dxc, x_local, x_c, st, join);
      h->norms_valid = h->auto_residual;
    }
    return (int)PP_SUCCESSFUL;
  });
}

static void enqueue_residual_local(pp_handle *h, double *buf_dev, cudaStream_t st) {
  if (h->res_blocks > 0) {
    residual_rows_kernel<<<h->res_blocks, 256, 0, st>>>(h->rl_ptr.p, h->rl_col.p, h->rl_src.p, h->last_vals, h->last_x,
                                                       h->last_xc, h->local_dim, h->last_rhs, h->res_loc.p,
                                                       h->res_part.p, h->local_dim);
    h->launches++;
  }
  residual_border_kernel<<<std::max(1, (h->m_c + 3) / 4), 128, 0, st>>>(   // a warp per coupling row
      h->rb_ptr.p, h->rb_col.p, h->rb_src.p, h->last_vals, h->last_x, h->m_c, h->res_part.p, h->res_blocks, buf_dev);
  h->launches++;
}

static void enqueue_residual_norms(pp_handle *h, const double *buf_sum_dev, cudaStream_t st) {
  residual_coupling_kernel<<<1, 256, 0, st>>>(h->rq_ptr.p, h->rq_col.p, h->rq_src.p, h->last_vals, h->last_xc,
                                             h->bc_keep.p, buf_sum_dev, h->m_c, h->res_c.p, h->res_out.p);
  h->launches++;
  h->pin_out2.ensure(2);
  CK(cudaMemcpyAsync(h->pin_out2.p, h->res_out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
}

int pp_residual_local(pp_handle *h, double *buf_dev, void *stream) {
  if (!h || !h->solved || !h->last_vals) return misuse("pp_residual_local: a completed solve is required first");
  if (!buf_dev) return misuse("pp_residual_local: null buffer");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    h->norms_valid = false;
    enqueue_residual_local(h, buf_dev, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return (int)PP_SUCCESSFUL;
  });
}

int pp_residual_norms(pp_handle *h, const double *buf_sum_dev, double out[2], void *stream) {
  if (!h || !h->solved) return misuse("pp_residual_norms: a completed solve is required first");
  if (!out) return misuse("pp_residual_norms: null argument");
  if (h->norms_valid) {  // formed by pp_solve_backward (option "auto_residual"): nothing to launch or wait for
    out[0] = h->pin_out2.p[0];
    out[1] = h->pin_out2.p[1];
    return PP_SUCCESSFUL;
  }
  if (!buf_sum_dev) return misuse("pp_residual_norms: null argument");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    enqueue_residual_norms(h, buf_sum_dev, st);
    CK(cudaStreamSynchronize(st));
    out[0] = h->pin_out2.p[0];
    out[1] = h->pin_out2.p[1];
    return (int)PP_SUCCESSFUL;
  });
}

int pp_refine_forward(pp_handle *h, double *rc_local_dev, void *stream) {
  if (!h || !h->solved) return misuse("pp_refine_forward: a completed solve and pp_residual_norms are required first");
  if (h->m_c > 0 && !rc_local_dev) return misuse("pp_refine_forward: null coupling buffer");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    run_forward(h, h->res_loc.p, rc_local_dev, (cudaStream_t)stream);
    return (int)PP_SUCCESSFUL;
  });
}

int pp_refine_backward(pp_handle *h, const double *rc_sum_dev, int on_device, double *x_local, double *x_c,
                       void *stream) {
  if (!h || !h->solved) return misuse("pp_refine_backward: a completed solve is required first");
  if (h->m_c > 0 && !rc_sum_dev) return misuse("pp_refine_backward: null coupling argument");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    h->norms_valid = false;
    run_backward(h, rc_sum_dev, h->res_c.p, h->dx_tmp.p, h->dxc_tmp.p, st);
    if (h->local_dim > 0) {
      axpy1_kernel<<<(unsigned)((h->local_dim + 255) / 256), 256, 0, st>>>(h->last_x, h->dx_tmp.p, h->local_dim);
      h->launches++;
    }
    if (h->m_c > 0) {
      axpy1_kernel<<<(h->m_c + 255) / 256, 256, 0, st>>>(h->last_xc, h->dxc_tmp.p, h->m_c);
      h->launches++;
    }
    CK(cudaGetLastError());
    if (!on_device) {
      if ((h->local_dim > 0 && !x_local) || (h->m_c > 0 && !x_c)) return misuse("pp_refine_backward: null output");
      copy_out(h, h->last_x, h->last_xc, x_local, x_c, st);
    }
    return (int)PP_SUCCESSFUL;
  });
}

int64_t pp_factor_bytes(const pp_handle *h) {
  int64_t tot = 0;
  for (; h; h = h->child) tot += h->bytes + (int64_t)h->sc_nnz * 32;
  return tot;
}
int64_t pp_local_dim(const pp_handle *h) { return h ? h->local_dim : 0; }
int64_t pp_kernel_launches(const pp_handle *h) {
  int64_t tot = 0;
  for (; h; h = h->child) tot += h->launches;
  return tot;
}

#ifdef PP_TRACE_SOLVE
extern "C" int pp_debug_pc_trace(long long *out, int reset) {
  if (cudaMemcpyFromSymbol(out, ppb::g_pc_trace, sizeof(long long) * 24) != cudaSuccess) return 3;
  long long z[24] = {0};
  if (reset) cudaMemcpyToSymbol(ppb::g_pc_trace, z, sizeof(z));
  return 0;
}
extern "C" int pp_debug_cs_trace(long long *out, int reset) {
  if (cudaMemcpyFromSymbol(out, ppb::g_cs_trace, sizeof(long long) * 8) != cudaSuccess) return 3;
  long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (reset) cudaMemcpyToSymbol(ppb::g_cs_trace, z, sizeof(z));
  return 0;
}
#endif
#ifdef PP_TRACE
extern "C" int pp_debug_trace_reset() {
  int z = 0;
  return cudaMemcpyToSymbol(ppb::g_tp, &z, sizeof(int)) == cudaSuccess ? 0 : 3;
}
extern "C" int pp_debug_trace(long long *out, int n) {
  return cudaMemcpyFromSymbol(out, ppb::g_trace, sizeof(long long) * (size_t)std::min(n, 2048)) == cudaSuccess ? 0 : 3;
}
#endif

int pp_stage_values(pp_handle *h, int64_t nseg, void *const *ptr, const int64_t *off, const int64_t *len,
                    void *staging, int threads, int chunks, void *stream) {
  if (!h || !h->have_symbolic) return misuse("pp_stage_values: symbolic factorization required first");
  if (nseg < 0 || (nseg > 0 && (!ptr || !off || !len)) || !staging) return misuse("pp_stage_values: null argument");
  int64_t total = 0;
  for (int64_t k = 0; k < nseg; ++k) {
    if (len[k] < 0 || off[k] != total || (len[k] > 0 && !ptr[k])) return misuse("pp_stage_values: segments must tile the buffer in order");
    total += len[k];
  }
  if (total != (int64_t)h->nvals * (int64_t)sizeof(double)) return misuse("pp_stage_values: size does not match the analysed pattern");
  if (!is_pinned_host(staging)) return misuse("pp_stage_values: the staging buffer must be pinned host memory");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    h->staged_from = nullptr;
    const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(chunks, total / (512 << 10)));
    int64_t k0 = 0;
    for (int c = 0; c < nch; ++c) {
      // segments [k0, k1): up to the c-th share of the bytes; their transfer overlaps the gather of the next share
      const int64_t goal = total * (c + 1) / nch;
      int64_t k1 = k0, hi = k0 < nseg ? off[k0] : total;
      while (k1 < nseg && (off[k1] + len[k1] <= goal || c == nch - 1)) { hi = off[k1] + len[k1]; ++k1; }
      if (k1 == k0) continue;
      const int64_t lo = off[k0];
      CopyPool::instance().run(k1 - k0, ptr + k0, off + k0, len + k0, static_cast<char *>(staging), true, threads);
      CK(cudaMemcpyAsync(reinterpret_cast<char *>(h->vals.p) + lo, static_cast<char *>(staging) + lo, (size_t)(hi - lo),
                         cudaMemcpyHostToDevice, st));
      k0 = k1;
    }
    h->staged_from = staging;
    return (int)PP_SUCCESSFUL;
  });
}

int pp_host_copy(int64_t nseg, void *const *ptr, const int64_t *off, const int64_t *len, void *staging,
                 int to_staging, int threads) {
  if (nseg < 0 || (nseg > 0 && (!ptr || !off || !len || !staging))) return misuse("pp_host_copy: null argument");
  for (int64_t k = 0; k < nseg; ++k)
    if (len[k] < 0 || off[k] < 0 || (len[k] > 0 && !ptr[k])) return misuse("pp_host_copy: bad segment");
  try {
    CopyPool::instance().run(nseg, ptr, off, len, static_cast<char *>(staging), to_staging != 0, threads);
  } catch (const std::exception &e) {
    g_error = e.what();
    return PP_ERROR;
  }
  return PP_SUCCESSFUL;
}

int pp_host_wake(int threads, int64_t spin_ns) {
  if (threads < 0 || spin_ns < 0) return misuse("pp_host_wake: negative argument");
  try {
    CopyPool::instance().wake(threads, spin_ns);
  } catch (const std::exception &e) {
    g_error = e.what();
    return PP_ERROR;
  }
  return PP_SUCCESSFUL;
}

int pp_host_equal(int64_t nseg, void *const *a, void *const *b, const int64_t *len, int threads, int *equal) {
  if (nseg < 0 || !equal || (nseg > 0 && (!a || !b || !len))) return misuse("pp_host_equal: null argument");
  for (int64_t k = 0; k < nseg; ++k)
    if (len[k] < 0 || (len[k] > 0 && (!a[k] || !b[k]))) return misuse("pp_host_equal: bad segment");
  try {
    *equal = CopyPool::instance().equal(nseg, a, b, len, threads) ? 1 : 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return PP_ERROR;
  }
  return PP_SUCCESSFUL;
}

int pp_peer_allreduce(int world, int rank, const void *const *bufs, void *const *signal_pads, int slot, uint32_t seq,
                      int64_t n, double *out_dev, void *stream) {
  if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !bufs || !signal_pads || n < 0 || slot < 0 ||
      (n > 0 && !out_dev) || n > (1 << 24))
    return misuse("pp_peer_allreduce: bad argument");
  return guarded([&]() {
    PeerPtrs P;
    for (int q = 0; q < PEER_MAX; ++q) {
      P.buf[q] = q < world ? static_cast<const double *>(bufs[q]) : nullptr;
      P.sig[q] = q < world ? static_cast<uint32_t *>(signal_pads[q]) : nullptr;
      if (q < world && (!P.buf[q] || !P.sig[q])) return misuse("pp_peer_allreduce: null peer pointer");
    }
    const int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(32, (n + 31) / 32 * 32));
    const int ctas = (int)std::min<int64_t>(32, std::max<int64_t>(1, (n + 2047) / 2048));   // about two elements per thread
    peer_allreduce_kernel<<<ctas, threads, 0, (cudaStream_t)stream>>>(P, rank, world, slot, seq, (int)n, out_dev);
    CK(cudaGetLastError());
    return (int)PP_SUCCESSFUL;
  });
}

int pp_profile(pp_handle *h, double *ms, int64_t *launches, int reset) {
  if (!h) return misuse("pp_profile: null handle");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    resolve_profile(h);
    for (int c = 0; c < PP_PROF_CLASSES; ++c) {
      if (ms) ms[c] = h->prof_ms[c];
      if (launches) launches[c] = h->prof_launches[c];
      if (reset) { h->prof_ms[c] = 0; h->prof_launches[c] = 0; }
    }
    return (int)PP_SUCCESSFUL;
  });
}

int pp_plan_stats(pp_handle *h, int32_t block, int64_t out[12]) {
  if (!h || !h->have_symbolic || block < 0 || block >= h->n_local || !out) return misuse("pp_plan_stats: bad argument");
  const PatternPlan &P = h->plans[h->block_plan[block]];
  out[0] = h->block_plan[block];
  out[1] = P.ns;
  out[2] = P.nT;
  out[3] = P.DR;
  out[4] = P.nnz_l;
  out[5] = P.max_front;
  out[6] = P.l_total;
  out[7] = P.stack_cap;
  out[8] = (int64_t)h->plans.size();
  out[9] = h->sparse_failed ? 1 : 0;
  out[10] = 0;
  out[11] = 0;
  if (P.ns > 0) {  // delayed pivots that reached the root in the last factorisation
    int info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<SparseBlock> sb(1);
    if (cudaMemcpy(sb.data(), h->blocks_dev.p + block, sizeof(SparseBlock), cudaMemcpyDeviceToHost) == cudaSuccess &&
        cudaMemcpy(info, sb[0].info, sizeof(info), cudaMemcpyDeviceToHost) == cudaSuccess) {
      out[10] = info[1];
      out[11] = info[0] + 10 * info[3] + 1000 * (int64_t)info[4] + 100000000ll * info[5];
    }
  }
  return PP_SUCCESSFUL;
}

// ---- host-only access to the symbolic analysis (tests, documentation of the plan format) ----
struct pp_plan {
  PatternPlan P;
};

static int g_plan_ordering = 0;
int pp_plan_set_ordering(int32_t ordering) {
  g_plan_ordering = ordering;
  return PP_SUCCESSFUL;
}

// representative values of the entries of the NEXT pp_plan_create (what pp_symbolic's `values_hint` is to a handle):
// consumed by that call, ignored if its length is not the entry count; NULL / 0 clears it
static std::vector<double> g_plan_hint;
int pp_plan_set_hint(const double *values, int64_t nent) {
  if (nent < 0 || (nent > 0 && !values)) return misuse("pp_plan_set_hint: bad argument");
  return guarded([&]() {
    g_plan_hint.assign(values, values + nent);
    return (int)PP_SUCCESSFUL;
  });
}

int pp_plan_create(int32_t n, int32_t m, int64_t nent, const int32_t *rows, const int32_t *cols, int32_t fmax,
                   int32_t dmax, int32_t min_sparse_n, pp_plan **out) {
  std::vector<double> hint;
  hint.swap(g_plan_hint);   // consumed, whatever happens below
  if (!out || n < 0 || m < 0 || nent < 0 || (nent > 0 && (!rows || !cols))) return misuse("pp_plan_create: bad argument");
  *out = nullptr;
  return guarded([&]() {
    std::vector<int> r(rows, rows + nent), c(cols, cols + nent), src((size_t)nent);
    std::iota(src.begin(), src.end(), 0);
    for (int64_t k = 0; k < nent; ++k)
      if (r[k] < c[k] || c[k] < 0 || c[k] >= n || r[k] >= n + m) return misuse("pp_plan_create: entry outside the lower triangle");
    PlanOptions opt;
    if (fmax > 0) opt.fmax = std::max(8, std::min((int)fmax, SF_SBUF - 8));
    if (dmax >= 0) opt.dmax = std::max(0, std::min((int)dmax, SF_SBUF / 2));
    if (min_sparse_n >= 0) opt.min_sparse_n = min_sparse_n;
    opt.ordering = g_plan_ordering;
    auto *pl = new pp_plan();
    pl->P = build_plan(n, m, r, c, src, opt, false, (nent > 0 && (int64_t)hint.size() == nent) ? &hint : nullptr);
    *out = pl;
    return (int)PP_SUCCESSFUL;
  });
}

int pp_plan_get(const pp_plan *pl, const char *name, const int32_t **data, int64_t *len) {
  if (!pl || !name || !data || !len) return misuse("pp_plan_get: null argument");
  const PatternPlan &P = pl->P;
  const std::string k(name);
  const std::vector<int> *v = nullptr;
  if (k == "rootcols") v = &P.rootcols;
  else if (k == "col_ptr") v = &P.col_ptr;
  else if (k == "cols") v = &P.cols;
  else if (k == "row_ptr") v = &P.row_ptr;
  else if (k == "rows") v = &P.rows;
  else if (k == "rel") v = &P.rel;
  else if (k == "parent") v = &P.parent;
  else if (k == "nchild") v = &P.nchild;
  else if (k == "dslot") v = &P.dslot;
  else if (k == "child_ptr") v = &P.child_ptr;
  else if (k == "child_idx") v = &P.child_idx;
  else if (k == "root_children") v = &P.root_children;
  else if (k == "tiny_ptr") v = &P.tiny_ptr;
  else if (k == "tiny_idx") v = &P.tiny_idx;
  else if (k == "med_ptr") v = &P.med_ptr;
  else if (k == "med_idx") v = &P.med_idx;
  else if (k == "big_ptr") v = &P.big_ptr;
  else if (k == "big_idx") v = &P.big_idx;
  else if (k == "dcap") v = &P.dcap;
  else if (k == "ent_ptr") v = &P.ent_ptr;
  else if (k == "tgt_row") v = &P.tgt_row;
  else if (k == "tgt_col") v = &P.tgt_col;
  else if (k == "tgt_src_ptr") v = &P.tgt_src_ptr;
  else if (k == "tgt_src") v = &P.tgt_src;
  else if (k == "root_row") v = &P.root_row;
  else if (k == "root_col") v = &P.root_col;
  else if (k == "root_src") v = &P.root_src;
  else return misuse("pp_plan_get: unknown array " + k);
  *data = v->data();
  *len = (int64_t)v->size();
  return PP_SUCCESSFUL;
}

int pp_plan_scalar(const pp_plan *pl, const char *name, int64_t *value) {
  if (!pl || !name || !value) return misuse("pp_plan_scalar: null argument");
  const PatternPlan &P = pl->P;
  const std::string k(name);
  if (k == "n") *value = P.n;
  else if (k == "m") *value = P.m;
  else if (k == "nT") *value = P.nT;
  else if (k == "DR") *value = P.DR;
  else if (k == "ns") *value = P.ns;
  else if (k == "nnz_l") *value = P.nnz_l;
  else if (k == "max_front") *value = P.max_front;
  else if (k == "l_total") *value = P.l_total;
  else if (k == "stack_cap") *value = P.stack_cap;
  else if (k == "nlevels") *value = P.nlevels;
  else return misuse("pp_plan_scalar: unknown scalar " + k);
  return PP_SUCCESSFUL;
}

int pp_plan_destroy(pp_plan *pl) {
  delete pl;
  return PP_SUCCESSFUL;
}

// ---- host-only access to the analysis of a sparse coupling system (tests; no GPU needed) ----
struct pp_cplan {
  CouplingLevel L;
};

int pp_cplan_create(int32_t m_c, int32_t n_cliques, const int64_t *ptr, const int32_t *rows, int64_t nq,
                    const int32_t *qrow, const int32_t *qcol, int32_t min_mc, double max_density, pp_cplan **out) {
  if (!out || m_c < 0 || n_cliques < 0 || nq < 0 || (n_cliques > 0 && !ptr) || (nq > 0 && (!qrow || !qcol)))
    return misuse("pp_cplan_create: bad argument");
  *out = nullptr;
  return guarded([&]() {
    std::vector<int64_t> cp(1, 0);
    std::vector<int32_t> cr;
    if (n_cliques > 0) {
      if (ptr[0] != 0) return misuse("pp_cplan_create: clique_ptr must start at 0");
      for (int k = 0; k < n_cliques; ++k)
        if (ptr[k + 1] < ptr[k]) return misuse("pp_cplan_create: clique_ptr must not decrease");
      if (ptr[n_cliques] > 0 && !rows) return misuse("pp_cplan_create: null clique rows");
      cp.assign(ptr, ptr + n_cliques + 1);
      cr.assign(rows, rows + ptr[n_cliques]);
    }
    for (int32_t r : cr)
      if (r < 0 || r >= m_c) return misuse("pp_cplan_create: clique row out of range");
    std::vector<int32_t> qr(qrow, qrow + nq), qc(qcol, qcol + nq);
    for (int64_t k = 0; k < nq; ++k)
      if (qc[(size_t)k] < 0 || qr[(size_t)k] < qc[(size_t)k] || qr[(size_t)k] >= m_c) return misuse("pp_cplan_create: Q entry outside the lower triangle");
    CouplingOptions opt;
    if (min_mc >= 0) opt.min_mc = min_mc;
    if (max_density >= 0.0) opt.max_density = max_density;
    auto *pl = new pp_cplan();
    pl->L = analyse_coupling(m_c, cp, cr, qr, qc, opt);
    *out = pl;
    return (int)PP_SUCCESSFUL;
  });
}

int pp_cplan_get(const pp_cplan *pl, const char *name, int64_t *buf, int64_t cap, int64_t *len) {
  if (!pl || !name || !len) return misuse("pp_cplan_get: null argument");
  const CouplingLevel &L = pl->L;
  const std::string k(name);
  std::vector<int64_t> v;
  auto widen = [&](const std::vector<int32_t> &a) { v.assign(a.begin(), a.end()); };
  if (k == "scalars") v = {L.sparse ? 1 : 0, L.n_blocks, L.m_next, L.nnz(), L.m_c};
  else if (k == "colptr") v = L.colptr;
  else if (k == "border_ptr") v = L.border_ptr;
  else if (k == "rowidx") widen(L.rowidx);
  else if (k == "block_n") widen(L.block_n);
  else if (k == "border_rows") widen(L.border_rows);
  else if (k == "dest_front") widen(L.dest_front);
  else if (k == "dest_row") widen(L.dest_row);
  else if (k == "dest_col") widen(L.dest_col);
  else if (k == "perm_local") widen(L.perm_local);
  else if (k == "perm_c") widen(L.perm_c);
  else return misuse("pp_cplan_get: unknown array " + k);
  *len = (int64_t)v.size();
  if (buf && cap >= (int64_t)v.size()) std::copy(v.begin(), v.end(), buf);
  return PP_SUCCESSFUL;
}

int pp_cplan_destroy(pp_cplan *pl) {
  delete pl;
  return PP_SUCCESSFUL;
}

int pp_debug_front(pp_handle *h, int32_t f, double *out, int64_t out_len, int32_t *ld, int32_t *piv,
                   int32_t *bsz) {
  if (!h || !h->have_symbolic || f < 0 || f > h->n_local) return misuse("pp_debug_front: bad front index");
  return guarded([&]() {
    DeviceGuard dev_guard(h->device);
    CK(cudaDeviceSynchronize());
    Front F;
    CK(cudaMemcpy(&F, h->fronts.p + f, sizeof(Front), cudaMemcpyDeviceToHost));
    const int64_t need = (int64_t)F.ld * F.nf;
    if (ld) *ld = F.ld;
    if (out) {
      if (out_len < need) return misuse("pp_debug_front: output buffer too small");
      CK(cudaMemcpy(out, F.A, (size_t)need * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (piv) CK(cudaMemcpy(piv, F.ipiv, (size_t)F.n * sizeof(int), cudaMemcpyDeviceToHost));
    if (bsz) CK(cudaMemcpy(bsz, F.bsz, (size_t)F.n * sizeof(int), cudaMemcpyDeviceToHost));
    return (int)PP_SUCCESSFUL;
  });
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Interior-point vector kernels (N3; csrc/ipmvec.cuh).  Stateless: device pointers in, results in a device buffer.
// ---------------------------------------------------------------------------------------------------------------
namespace {
int g_iv_sms = 0;
int iv_grid(int64_t n) {
  if (!g_iv_sms) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    g_iv_sms = sms;
  }
  // whole multiples of the SM count, eight CTAs of 256 threads per SM at most, one element pair per thread at least
  const int64_t want = (n / 2 + IV_NT - 1) / IV_NT;
  int64_t per_sm = (want + g_iv_sms - 1) / g_iv_sms;
  per_sm = std::max<int64_t>(1, std::min<int64_t>(per_sm, IV_MAXCTA / 148));
  return (int)std::min<int64_t>(per_sm * g_iv_sms, IV_MAXCTA);
}
bool iv_aligned(std::initializer_list<const void *> ps) {
  for (const void *p : ps)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15u)) return false;
  return true;
}
double *iv_ws(void *workspace) { return reinterpret_cast<double *>(workspace) + 2; }
unsigned int *iv_counter(void *workspace) { return reinterpret_cast<unsigned int *>(workspace); }
int iv_done(const char *what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
  return PP_SUCCESSFUL;
}
}  // namespace

extern "C" {

int64_t pp_ipm_workspace_bytes(void) { return (int64_t)(2 + (size_t)IV_MAXCTA * IV_SLOTS) * sizeof(double); }

int pp_ipm_fill(double *out, int32_t n, double value, void *stream) {
  if (!out || n < 0 || n > 1024) return misuse("pp_ipm_fill: bad arguments");
  if (n) ipm_fill_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(out, n, value);
  return iv_done("pp_ipm_fill");
}

int pp_ipm_fraction_to_boundary(int64_t n, double tau, double barrier, const double *x, const double *dx,
                                const double *lb, const double *ub, const double *zl, const double *zu,
                                double *out2, void *workspace, void *stream) {
  if (n < 0 || !out2 || !workspace || (n && (!x || !dx || !lb || !ub || !zl || !zu)))
    return misuse("pp_ipm_fraction_to_boundary: null pointer");
  if (n == 0) return PP_SUCCESSFUL;   // an empty group leaves the running minima as they are (:660-661)
  ipm_ftb_kernel<<<iv_grid(n), IV_NT, 0, (cudaStream_t)stream>>>(n, tau, barrier, x, dx, lb, ub, zl, zu, iv_ws(workspace),
                                                                 iv_counter(workspace), out2,
                                                                 iv_aligned({x, dx, lb, ub, zl, zu}) ? 1 : 0);
  return iv_done("pp_ipm_fraction_to_boundary");
}

int pp_ipm_complementarity(int64_t n, double barrier, const double *x, const double *lb, const double *ub,
                           const double *zl, const double *zu, double *out6, void *workspace, void *stream) {
  if (n < 0 || !out6 || !workspace || (n && (!x || !lb || !ub || !zl || !zu)))
    return misuse("pp_ipm_complementarity: null pointer");
  if (n == 0) return PP_SUCCESSFUL;
  ipm_compl_kernel<<<iv_grid(n), IV_NT, 0, (cudaStream_t)stream>>>(n, barrier, x, lb, ub, zl, zu, iv_ws(workspace),
                                                                   iv_counter(workspace), out6,
                                                                   iv_aligned({x, lb, ub, zl, zu}) ? 1 : 0);
  return iv_done("pp_ipm_complementarity");
}

int pp_ipm_max_abs(int64_t n, const double *a, const double *b, double *out2, void *workspace, void *stream) {
  if (n < 0 || !out2 || !workspace || (n && !a)) return misuse("pp_ipm_max_abs: null pointer");
  if (n == 0) return PP_SUCCESSFUL;
  ipm_maxabs_kernel<<<iv_grid(n), IV_NT, 0, (cudaStream_t)stream>>>(n, a, b, iv_ws(workspace), iv_counter(workspace), out2,
                                                                    iv_aligned({a, b}) ? 1 : 0);
  return iv_done("pp_ipm_max_abs");
}

int pp_ipm_step(int64_t n, const double *alpha3, double barrier, double *x, const double *dx, const double *lb,
                const double *ub, double *zl, double *zu, void *stream) {
  if (n < 0 || !alpha3 || (n && (!x || !dx || !lb || !ub || !zl || !zu))) return misuse("pp_ipm_step: null pointer");
  if (n == 0) return PP_SUCCESSFUL;
  ipm_step_kernel<<<iv_grid(2 * n), IV_NT, 0, (cudaStream_t)stream>>>(n, alpha3, barrier, x, dx, lb, ub, zl, zu);
  return iv_done("pp_ipm_step");
}

int pp_ipm_axpy(int64_t n, const double *alpha3, int32_t which, double *y, const double *dy, void *stream) {
  if (n < 0 || !alpha3 || which < 0 || which > 1 || (n && (!y || !dy))) return misuse("pp_ipm_axpy: bad arguments");
  if (n == 0) return PP_SUCCESSFUL;
  ipm_axpy_kernel<<<iv_grid(2 * n), IV_NT, 0, (cudaStream_t)stream>>>(n, alpha3, which, y, dy);
  return iv_done("pp_ipm_axpy");
}

}  // extern "C"
