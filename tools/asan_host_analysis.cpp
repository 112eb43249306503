// Sanitizer harness for the host analyses (csrc/symbolic.hpp, csrc/coupling.hpp): random structured patterns, with and
// without a values hint, every ordering, small and large front caps -- compiled with AddressSanitizer, UBSan and the
// bounds-checked libstdc++ (`-D_GLIBCXX_ASSERTIONS`), so that an out-of-range index or an overflow inside the analysis
// aborts instead of passing silently.  Build + run (no GPU, no CUDA):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -D_GLIBCXX_ASSERTIONS \
//       -I parapint_b200/csrc tools/asan_host_analysis.cpp -o /tmp/asan_host_analysis && /tmp/asan_host_analysis 400 [seed]
#include <cstdio>
#include <cstdlib>
#include <random>

#include "coupling.hpp"
#include "symbolic.hpp"

using namespace ppb;

static std::mt19937_64 rng(12345);
static int ri(int lo, int hi) { return (int)(lo + rng() % (uint64_t)(hi - lo + 1)); }  // inclusive
static double rd() { return (double)(rng() >> 11) / 9007199254740992.0; }

struct Pattern {
  int n = 0, m = 0;
  std::vector<int> rows, cols, src;
  std::vector<double> vals;
  void add(int r, int c, double v) {
    if (r < c) std::swap(r, c);
    rows.push_back(r); cols.push_back(c); src.push_back((int)src.size()); vals.push_back(v);
  }
};

static Pattern make(int kind) {
  Pattern P;
  const int n = P.n = ri(1, 500);
  auto diag = [&](int from, int to, bool zero_some) {
    for (int i = from; i < to; ++i)
      if (!zero_some || rd() < 0.5) P.add(i, i, zero_some && rd() < 0.5 ? 0.0 : 1.0 + rd());
  };
  switch (kind) {
    case 0:  // random
      diag(0, n, false);
      for (int k = 0, e = ri(0, 4 * n); k < e; ++k) P.add(ri(0, n - 1), ri(0, n - 1), rd() - 0.5);
      break;
    case 1: {  // band
      diag(0, n, false);
      const int w = ri(1, 6);
      for (int i = 0; i < n; ++i)
        for (int d = 1; d <= w && i + d < n; ++d) P.add(i + d, i, rd() - 0.5);
      break;
    }
    case 2: {  // arrow: a few hubs
      diag(0, n, false);
      for (int h = 0, nh = ri(1, 3); h < nh; ++h) {
        const int hub = ri(0, n - 1);
        for (int i = 0; i < n; ++i)
          if (rd() < 0.4) P.add(hub, i, rd() - 0.5);
      }
      break;
    }
    case 3: {  // KKT: multiplier columns with no (or a numerically zero) diagonal
      const int nx = std::max(1, (int)(n * (0.5 + 0.3 * rd())));
      diag(0, nx, false);
      diag(nx, n, true);
      for (int k = 0, e = ri(0, 2 * nx); k < e; ++k) P.add(ri(0, nx - 1), ri(0, nx - 1), rd() - 0.5);
      for (int r = nx; r < n; ++r)
        for (int k = 0, e = ri(0, 3); k < e; ++k) P.add(r, ri(0, nx - 1), rd() < 0.2 ? 0.0 : rd() - 0.5);
      break;
    }
    default:  // nothing but (some) diagonal entries, duplicates
      diag(0, n, true);
      diag(0, n, true);
  }
  P.m = rd() < 0.3 ? 0 : ri(1, 40);
  for (int a = 0; a < P.m; ++a)       // every border row has an entry (rows without one are not border rows)
    for (int k = 0, e = ri(1, 4); k < e; ++k) {
      P.rows.push_back(n + a); P.cols.push_back(ri(0, n - 1)); P.src.push_back((int)P.src.size()); P.vals.push_back(rd() - 0.5);
    }
  return P;
}

static void check_plan(const PatternPlan &pl, const Pattern &P) {
  std::vector<int> seen(P.n, 0);
  for (int c : pl.rootcols) seen.at(c)++;
  for (int c : pl.cols) seen.at(c)++;
  for (int v : seen)
    if (v != 1) { std::fprintf(stderr, "column eliminated %d times\n", v); std::abort(); }
  for (int s = 0; s < pl.ns; ++s) {
    const int p = pl.parent.at(s);
    if (p != -1 && p <= s) { std::fprintf(stderr, "not a postorder\n"); std::abort(); }
  }
}

int main(int argc, char **argv) {
  const int cases = argc > 1 ? std::atoi(argv[1]) : 200;
  if (argc > 2) rng.seed((uint64_t)std::atoll(argv[2]));
  for (int c = 0; c < cases; ++c) {
    Pattern P = make(c % 5);
    PlanOptions opt;
    opt.ordering = ri(0, 3);
    opt.min_sparse_n = (c & 1) ? 16 : 64;
    if (rd() < 0.3) opt.fmax = 32;
    if (rd() < 0.3) opt.dmax = ri(0, 48);
    if (rd() < 0.2) opt.pair_weak = false;
    const bool hinted = rd() < 0.5;
    PatternPlan pl = build_plan(P.n, P.m, P.rows, P.cols, P.src, opt, rd() < 0.05, hinted ? &P.vals : nullptr);
    check_plan(pl, P);
    // coupling analysis on random cliques
    const int m_c = ri(0, 300);
    std::vector<int64_t> ptr(1, 0);
    std::vector<int32_t> rows, qr, qc;
    if (m_c > 0) {
      const int style = c % 3;
      for (int k = 0, e = ri(0, 40); k < e; ++k) {
        std::vector<int32_t> cl;
        if (style == 0) {        // chain-like
          const int g = std::max(1, m_c / 40), t = ri(0, 39);
          for (int v = t * g; v < std::min(m_c, (t + 2) * g); ++v) cl.push_back(v);
        } else {
          for (int v = 0; v < m_c; ++v)
            if (rd() < (style == 1 ? 0.03 : 0.2)) cl.push_back(v);
        }
        rows.insert(rows.end(), cl.begin(), cl.end());
        ptr.push_back((int64_t)rows.size());
      }
      for (int i = 0; i < m_c; ++i)
        if (rd() < 0.9) { qr.push_back(i); qc.push_back(i); }
      for (int k = 0, e = ri(0, 20); k < e; ++k) {
        int a = ri(0, m_c - 1), b = ri(0, m_c - 1);
        qr.push_back(std::max(a, b)); qc.push_back(std::min(a, b));
      }
    }
    CouplingOptions co;
    co.min_mc = ri(2, 64);
    co.max_density = 0.2 + 0.7 * rd();
    CouplingLevel L = analyse_coupling(m_c, ptr, rows, qr, qc, co);
    if (L.sparse) {
      std::vector<int> cnt(m_c, 0);
      for (int v : L.perm_local) cnt.at(v)++;
      for (int v : L.perm_c) cnt.at(v)++;
      for (int v : cnt)
        if (v != 1) { std::fprintf(stderr, "coupling variable placed %d times\n", v); std::abort(); }
      if ((int64_t)L.dest_front.size() != L.nnz()) { std::fprintf(stderr, "dest size\n"); std::abort(); }
    }
    if (c % 50 == 0) std::printf("case %d: n %d m %d ns %d nT %d | m_c %d sparse %d blocks %d\n", c, P.n, P.m, pl.ns, pl.nT, m_c, (int)L.sparse, L.n_blocks);
  }
  std::printf("ok: %d cases\n", cases);
  return 0;
}
