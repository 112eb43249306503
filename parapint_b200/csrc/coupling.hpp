// Host-side analysis of a SPARSE Schur complement (the coupling system of time-decomposed problems).
//
// Reference: MPISchurComplementLinearSolver._get_sc_structure / _get_all_nonzero_elements_in_sc
// (parapint/linalg/schur_complement/mpi_explicit_schur_complement.py:88-125,228-255) -- the pattern of
// S = Q - sum_i A_i K_i^-1 A_i^T is the union, over ALL blocks, of nonzero_rows(A_i) x nonzero_rows(A_i) and of the
// pattern of Q; the reference then hands that sparse S to a sparse leaf solver (:352-360).
//
// Here the sparse S is factorised by the machinery that factorises the KKT matrix itself: S is viewed as a
// block-bordered matrix once more.  Its variables are grouped by the set of cliques (block borders) they belong to;
// an independent set of groups (no S entry between two of them) becomes the diagonal blocks of the next level,
// the other variables its coupling system, whose Schur complement has again a pattern of cliques + Q -- and so on
// until the coupling system is small or dense.  For the block-tridiagonal S of a chain of time blocks
// (interfaces/schur_complement/sc_ip_interface.py:274-357) every level halves the chain: block cyclic reduction,
// log2(N) levels, each one batched launch sequence.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <numeric>
#include <vector>

namespace ppb {

struct CouplingOptions {
  int min_mc = 384;          // smaller coupling systems are dense fronts
  double max_density = 0.30; // ... as are those whose pattern fills more than this share of the triangle
  int max_levels = 40;
};

struct CouplingLevel {
  int m_c = 0;
  bool sparse = false;               // false: keep S dense (nothing below is filled in)
  // pattern of S, lower triangle by columns (rows ascending inside a column)
  std::vector<int64_t> colptr;       // [m_c + 1]
  std::vector<int32_t> rowidx;       // [nnz]
  // the next level: S as a block-bordered matrix
  int n_blocks = 0, m_next = 0;
  std::vector<int32_t> block_n;      // [n_blocks]
  std::vector<int64_t> border_ptr;   // [n_blocks + 1]
  std::vector<int32_t> border_rows;  // ascending indices into the next coupling system
  std::vector<int32_t> dest_front, dest_row, dest_col;  // one per pattern entry (pattern order)
  std::vector<int32_t> perm_local;   // variable of this level at position j of the next level's local vector
  std::vector<int32_t> perm_c;       // variable of this level that is variable j of the next coupling system
  int64_t nnz() const { return (int64_t)rowidx.size(); }
};

// cliques: sorted ascending row lists (one per block of ANY rank), q*: entries of Q with qrow >= qcol.
inline CouplingLevel analyse_coupling(int m_c, const std::vector<int64_t> &clq_ptr, const std::vector<int32_t> &clq_rows,
                                      const std::vector<int32_t> &qrow, const std::vector<int32_t> &qcol,
                                      const CouplingOptions &opt) {
  CouplingLevel L;
  L.m_c = m_c;
  const int ncl = (int)clq_ptr.size() - 1;
  if (m_c < opt.min_mc || ncl <= 0) return L;
  // distinct cliques (scenarios of a stochastic programme all touch the same rows: one clique)
  std::map<std::vector<int32_t>, int> distinct;
  std::vector<const int32_t *> cbeg;
  std::vector<int> clen;
  double bound = 0.0;
  for (int k = 0; k < ncl; ++k) {
    const int64_t a = clq_ptr[k], b = clq_ptr[k + 1];
    if (b <= a) continue;
    std::vector<int32_t> key(clq_rows.begin() + a, clq_rows.begin() + b);
    if (distinct.emplace(std::move(key), (int)distinct.size()).second) {
      cbeg.push_back(clq_rows.data() + a);
      clen.push_back((int)(b - a));
      bound += 0.5 * (double)(b - a) * (double)(b - a + 1);
    }
  }
  const double tri = 0.5 * (double)m_c * (double)(m_c + 1);
  if (bound + (double)qrow.size() > opt.max_density * tri) return L;  // (upper bound on the pattern size)

  // ---- pattern: union of the clique squares and of Q ----
  std::vector<std::vector<int32_t>> col((size_t)m_c);
  for (size_t k = 0; k < cbeg.size(); ++k) {
    const int32_t *r = cbeg[k];
    for (int b = 0; b < clen[k]; ++b) {
      auto &c = col[(size_t)r[b]];
      c.insert(c.end(), r + b, r + clen[k]);  // rows >= r[b] of the clique (ascending)
    }
  }
  for (size_t k = 0; k < qrow.size(); ++k) col[(size_t)qcol[k]].push_back(qrow[k]);
  L.colptr.assign((size_t)m_c + 1, 0);
  for (int c = 0; c < m_c; ++c) {
    auto &v = col[(size_t)c];
    v.push_back(c);  // the diagonal is always present (pivots, regularisation)
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    L.colptr[(size_t)c + 1] = L.colptr[(size_t)c] + (int64_t)v.size();
  }
  if ((double)L.colptr[(size_t)m_c] > opt.max_density * tri) { L.colptr.clear(); return L; }
  L.rowidx.reserve((size_t)L.colptr[(size_t)m_c]);
  for (int c = 0; c < m_c; ++c) L.rowidx.insert(L.rowidx.end(), col[(size_t)c].begin(), col[(size_t)c].end());
  col.clear();
  col.shrink_to_fit();

  // ---- groups: variables with the same clique membership ----
  std::vector<std::vector<int>> member((size_t)m_c);
  {
    int id = 0;
    for (size_t k = 0; k < cbeg.size(); ++k, ++id)
      for (int b = 0; b < clen[k]; ++b) member[(size_t)cbeg[k][b]].push_back(id);
  }
  std::vector<int> grp((size_t)m_c, -1);
  std::vector<int> gsize;
  {
    std::map<std::vector<int>, int> seen;
    for (int v = 0; v < m_c; ++v) {
      if (member[(size_t)v].empty()) {  // touched by Q only: a group of its own
        grp[(size_t)v] = (int)gsize.size();
        gsize.push_back(1);
        continue;
      }
      auto it = seen.find(member[(size_t)v]);
      if (it == seen.end()) {
        it = seen.emplace(member[(size_t)v], (int)gsize.size()).first;
        gsize.push_back(0);
      }
      grp[(size_t)v] = it->second;
      gsize[(size_t)it->second]++;
    }
  }
  const int ng = (int)gsize.size();
  // group graph from the pattern
  std::vector<std::vector<int>> gadj((size_t)ng);
  for (int c = 0; c < m_c; ++c)
    for (int64_t p = L.colptr[(size_t)c]; p < L.colptr[(size_t)c + 1]; ++p) {
      const int a = grp[(size_t)c], b = grp[(size_t)L.rowidx[(size_t)p]];
      if (a != b) { gadj[(size_t)a].push_back(b); gadj[(size_t)b].push_back(a); }
    }
  for (auto &a : gadj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  // greedy independent set, fewest neighbours first (ties: lowest group index) -- on a chain: every other group
  std::vector<int> order((size_t)ng);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return gadj[(size_t)a].size() < gadj[(size_t)b].size(); });
  std::vector<char> state((size_t)ng, 0);  // 1 = selected (a block of the next level), 2 = excluded
  for (int g : order) {
    if (state[(size_t)g]) continue;
    state[(size_t)g] = 1;
    for (int o : gadj[(size_t)g]) if (!state[(size_t)o]) state[(size_t)o] = 2;
  }
  // blocks in order of their first variable; variables keep their relative order everywhere
  std::vector<int> block_of_group((size_t)ng, -1);
  std::vector<int> pos((size_t)m_c, -1);   // position inside its block, or index in the next coupling system
  for (int v = 0; v < m_c; ++v) {
    const int g = grp[(size_t)v];
    if (state[(size_t)g] == 1) {
      if (block_of_group[(size_t)g] < 0) {
        block_of_group[(size_t)g] = L.n_blocks++;
        L.block_n.push_back(0);
      }
      pos[(size_t)v] = L.block_n[(size_t)block_of_group[(size_t)g]]++;
    } else {
      pos[(size_t)v] = L.m_next++;
      L.perm_c.push_back(v);
    }
  }
  if (L.n_blocks == 0 || L.m_next == m_c) { L.colptr.clear(); L.rowidx.clear(); return L; }
  std::vector<int64_t> boff((size_t)L.n_blocks + 1, 0);
  for (int b = 0; b < L.n_blocks; ++b) boff[(size_t)b + 1] = boff[(size_t)b] + L.block_n[(size_t)b];
  L.perm_local.assign((size_t)boff[(size_t)L.n_blocks], 0);
  for (int v = 0; v < m_c; ++v) {
    const int g = grp[(size_t)v];
    if (state[(size_t)g] == 1) L.perm_local[(size_t)(boff[(size_t)block_of_group[(size_t)g]] + pos[(size_t)v])] = v;
  }
  // borders: coupling variables of the next level with an entry against the block
  std::vector<std::vector<int32_t>> brows((size_t)L.n_blocks);
  for (int c = 0; c < m_c; ++c)
    for (int64_t p = L.colptr[(size_t)c]; p < L.colptr[(size_t)c + 1]; ++p) {
      const int r = L.rowidx[(size_t)p];
      const bool cs = state[(size_t)grp[(size_t)c]] == 1, rs = state[(size_t)grp[(size_t)r]] == 1;
      if (cs && !rs) brows[(size_t)block_of_group[(size_t)grp[(size_t)c]]].push_back(pos[(size_t)r]);
      else if (rs && !cs) brows[(size_t)block_of_group[(size_t)grp[(size_t)r]]].push_back(pos[(size_t)c]);
    }
  L.border_ptr.assign((size_t)L.n_blocks + 1, 0);
  for (int b = 0; b < L.n_blocks; ++b) {
    auto &v = brows[(size_t)b];
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    L.border_ptr[(size_t)b + 1] = L.border_ptr[(size_t)b] + (int64_t)v.size();
    L.border_rows.insert(L.border_rows.end(), v.begin(), v.end());
  }
  // destination of every pattern entry in the next level's fronts
  const int64_t nnz = L.nnz();
  L.dest_front.assign((size_t)nnz, -1);
  L.dest_row.assign((size_t)nnz, 0);
  L.dest_col.assign((size_t)nnz, 0);
  auto border_pos = [&](int b, int cvar) {
    const auto &v = brows[(size_t)b];
    return (int)(std::lower_bound(v.begin(), v.end(), cvar) - v.begin());
  };
  for (int c = 0; c < m_c; ++c)
    for (int64_t p = L.colptr[(size_t)c]; p < L.colptr[(size_t)c + 1]; ++p) {
      const int r = L.rowidx[(size_t)p];
      const int gc = grp[(size_t)c], gr = grp[(size_t)r];
      const bool cs = state[(size_t)gc] == 1, rs = state[(size_t)gr] == 1;
      if (cs && rs) {          // same block (two selected groups never share an entry)
        const int b = block_of_group[(size_t)gc];
        L.dest_front[(size_t)p] = b;
        L.dest_row[(size_t)p] = pos[(size_t)r];   // r >= c and positions follow the variable order
        L.dest_col[(size_t)p] = pos[(size_t)c];
      } else if (cs) {         // column in a block, row in the next coupling system
        const int b = block_of_group[(size_t)gc];
        L.dest_front[(size_t)p] = b;
        L.dest_row[(size_t)p] = L.block_n[(size_t)b] + border_pos(b, pos[(size_t)r]);
        L.dest_col[(size_t)p] = pos[(size_t)c];
      } else if (rs) {         // row in a block, column in the next coupling system: the transposed border entry
        const int b = block_of_group[(size_t)gr];
        L.dest_front[(size_t)p] = b;
        L.dest_row[(size_t)p] = L.block_n[(size_t)b] + border_pos(b, pos[(size_t)c]);
        L.dest_col[(size_t)p] = pos[(size_t)r];
      } else {                 // both stay: an entry of the next level's Q
        L.dest_front[(size_t)p] = L.n_blocks;
        L.dest_row[(size_t)p] = pos[(size_t)r];
        L.dest_col[(size_t)p] = pos[(size_t)c];
      }
    }
  L.sparse = true;
  return L;
}

}  // namespace ppb
