// Host-side symbolic analysis of one diagonal KKT block for the multifrontal path.
//
// The reference leaves do this inside SuperLU / MA27 / MUMPS (COLAMD / AMD + elimination tree,
// scipy_interface.py:30, ma27_interface.py:79, mumps_interface.py:59).  Here: a quotient-graph
// minimum-degree ordering of K_i (columns touched by the border A_i are kept for last), relaxed
// supernode amalgamation, and a split of the assembly tree into
//   * SUBTREE supernodes: small fronts (<= fmax rows) that one CTA factors in shared memory, and
//   * the ROOT: every remaining column (ancestor-closed top of the tree, incl. the border-touched
//     columns), handled as ONE dense front [root cols | delayed-pivot slots | border rows] by the
//     batched Bunch-Kaufman kernels of factor.cuh.
// All index lists use ORIGINAL column numbers of the block.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <set>
#include <vector>

namespace ppb {

struct PatternPlan {
  int n = 0, m = 0;     // block order, number of (nonzero) border rows
  int nT = 0, DR = 0;   // root front: nT static columns, DR delayed-pivot slots
  int ns = 0;           // subtree supernodes (postorder)
  std::vector<int> rootcols;          // [nT] original ids in root order
  std::vector<int> col_ptr, cols;     // own columns per supernode
  std::vector<int> row_ptr, rows;     // static contribution rows per supernode (original ids)
  std::vector<int> rel;               // aligned with rows: position in the parent's [cols|rows] list, or in rootcols
  std::vector<int> parent, nchild, dcap;
  std::vector<int64_t> l_off;         // offset of the supernode's L block in the per-block factor arena
  std::vector<int> fid_off, fs_off;   // offsets of the id / pivot-flag lists
  int64_t l_total = 0;
  int fid_total = 0, fs_total = 0;
  int64_t stack_cap = 0;              // (unused by the level-scheduled kernels; kept for statistics)
  std::vector<int> child_ptr, child_idx;   // children of every supernode, ascending
  std::vector<int> root_children;          // supernodes whose parent is the dense root, ascending
  std::vector<int> dslot;                  // delayed columns a supernode may pass up
  std::vector<int64_t> cb_off;             // contribution slot (dim <= ncb + dslot, stored dim x dim)
  std::vector<int> vec_off;                // forward-solve contribution vector slot
  int64_t cb_total = 0;
  int vec_total = 0;
  int nlevels = 0;                         // height classes: level 0 = leaves
  std::vector<int> tiny_ptr, tiny_idx;     // per level: fronts small enough for one warp
  std::vector<int> med_ptr, med_idx;       // per level: fronts for a two-warp group (<= medium rows, any children)
  std::vector<int> big_ptr, big_idx;       // per level: fronts factored by the whole CTA
  int max_front = 0;                  // largest subtree front incl. delayed capacity actually allowed
  int64_t nnz_l = 0;                  // static entries of L in the subtree part (statistics)
  // original entries grouped by destination: unique targets with their sources (relative value index)
  std::vector<int> ent_ptr;           // [ns+1] ranges into tgt_*
  std::vector<int> tgt_row, tgt_col;  // local (row >= nc means contribution row index + nc)
  std::vector<int> tgt_src_ptr;       // [ntargets+1] into tgt_src
  std::vector<int> tgt_src;           // relative value indices
  // root entries (dense front positions; border rows already offset by nT + DR)
  std::vector<int> root_row, root_col, root_src;  // one per input entry kept
};

namespace detail {

// Minimum-degree ordering on a quotient graph with element absorption and exact external degrees.
// Vertices with hold[v] != 0 are never eliminated.  Returns the elimination order of the others and,
// for each eliminated vertex, its column structure (the variables adjacent at elimination time).
// `stage` (optional) constrains the order: all vertices of stage k are eliminated before any of stage
// k+1, minimum degree decides inside a stage (nested dissection supplies the stages).
inline void minimum_degree(int n, const std::vector<std::vector<int>> &adj, const std::vector<char> &hold,
                           const std::vector<int> &stage, std::vector<int> &order,
                           std::vector<std::vector<int>> &lstruct) {
  std::vector<std::vector<int>> vadj(adj), eadj(n), evars(n);
  std::vector<char> eliminated(n, 0), ealive(n, 0);
  std::vector<int> mark(n, -1), degree(n, 0);
  int stamp = 0;
  const bool staged = !stage.empty();
  const long long SM = 1ll << 32;  // key = stage * 2^32 + degree
  auto key = [&](int v, int d) { return (staged ? (long long)stage[v] * SM : 0ll) + d; };
  std::set<std::pair<long long, int>> heap;
  for (int v = 0; v < n; ++v) {
    degree[v] = (int)vadj[v].size();
    if (!hold[v]) heap.insert({key(v, degree[v]), v});
  }
  lstruct.assign(n, {});
  order.clear();
  std::vector<int> reach;
  while (!heap.empty()) {
    const int p = heap.begin()->second;
    heap.erase(heap.begin());
    // reach set of p
    ++stamp;
    mark[p] = stamp;
    reach.clear();
    for (int v : vadj[p])
      if (!eliminated[v] && mark[v] != stamp) { mark[v] = stamp; reach.push_back(v); }
    for (int e : eadj[p])
      if (ealive[e])
        for (int v : evars[e])
          if (!eliminated[v] && mark[v] != stamp) { mark[v] = stamp; reach.push_back(v); }
    std::sort(reach.begin(), reach.end());
    lstruct[p] = reach;
    eliminated[p] = 1;
    order.push_back(p);
    for (int e : eadj[p]) ealive[e] = 0;  // absorbed into the new element p
    ealive[p] = 1;
    evars[p] = reach;
    const int pstamp = stamp;  // members of the new element carry mark == pstamp
    for (int v : reach) {
      // prune variable edges now covered by element p, and dead elements
      auto &va = vadj[v];
      va.erase(std::remove_if(va.begin(), va.end(), [&](int w) { return w == p || mark[w] == pstamp || eliminated[w]; }),
               va.end());
      auto &ea = eadj[v];
      ea.erase(std::remove_if(ea.begin(), ea.end(), [&](int e) { return !ealive[e]; }), ea.end());
      ea.push_back(p);
    }
    for (int v : reach) {
      // exact external degree of v
      ++stamp;
      mark[v] = stamp;
      int d = 0;
      for (int w : vadj[v])
        if (mark[w] != stamp) { mark[w] = stamp; ++d; }
      for (int e : eadj[v])
        for (int w : evars[e])
          if (!eliminated[w] && mark[w] != stamp) { mark[w] = stamp; ++d; }
      if (!hold[v]) {
        heap.erase({key(v, degree[v]), v});
        heap.insert({key(v, d), v});
      }
      degree[v] = d;
    }
    // restore the element-membership stamp for the next prune (marks were overwritten above)
    stamp += 1;
  }
}

// Nested dissection by breadth-first level structures (George): recursively split every connected
// component at the level set nearest its middle, thinned to the vertices that touch the next level.
// Produces a stage number per vertex: the two halves (recursively) come before their separator.
// Held vertices are not part of the graph.  Banded / time-like structures, which minimum degree turns
// into one long dependency chain, become shallow trees.
inline void nested_dissection_stages(int n, const std::vector<std::vector<int>> &adj,
                                     const std::vector<char> &hold, int leaf_size, std::vector<int> &stage) {
  stage.assign(n, -1);
  std::vector<int> region(n, -1);  // current region id of every still-unassigned vertex
  for (int v = 0; v < n; ++v) region[v] = hold[v] ? -2 : 0;
  int next_region = 1, next_stage = 0;
  std::vector<int> dist(n, -1), queue;
  struct Task { std::vector<int> verts; int id; bool emit_only; };
  std::vector<Task> stack;
  {
    Task t0;
    t0.id = 0;
    t0.emit_only = false;
    for (int v = 0; v < n; ++v)
      if (!hold[v]) t0.verts.push_back(v);
    if (t0.verts.empty()) return;
    stack.push_back(std::move(t0));
  }
  auto bfs = [&](int start, int id, std::vector<int> &visited) {
    visited.clear();
    visited.push_back(start);
    dist[start] = 0;
    for (size_t h = 0; h < visited.size(); ++h) {
      const int v = visited[h];
      for (int w : adj[v])
        if (region[w] == id && dist[w] < 0) { dist[w] = dist[v] + 1; visited.push_back(w); }
    }
  };
  while (!stack.empty()) {
    Task t = std::move(stack.back());
    stack.pop_back();
    if (t.emit_only || (int)t.verts.size() <= leaf_size) {
      for (int v : t.verts) { stage[v] = next_stage; region[v] = -3; }
      ++next_stage;
      continue;
    }
    // connected component of the first vertex
    std::vector<int> comp;
    bfs(t.verts[0], t.id, comp);
    if (comp.size() < t.verts.size()) {
      // several components: peel this one off, keep the rest as a task of its own
      const int cid = next_region++;
      Task rest;
      rest.id = t.id;
      rest.emit_only = false;
      for (int v : comp) { region[v] = cid; dist[v] = -1; }
      for (int v : t.verts)
        if (region[v] == t.id) rest.verts.push_back(v);
      Task c;
      c.id = cid;
      c.emit_only = false;
      c.verts = comp;
      stack.push_back(std::move(rest));
      stack.push_back(std::move(c));
      continue;
    }
    // pseudo-peripheral start: repeat BFS from the farthest vertex a few times
    int start = comp.back();
    std::vector<int> lev;
    for (int it = 0; it < 3; ++it) {
      for (int v : comp) dist[v] = -1;
      bfs(start, t.id, lev);
      const int far = lev.back();
      if (it < 2 && dist[far] > 0) start = far;
    }
    const int depth = dist[lev.back()];
    if (depth < 2) {  // no useful separator: one leaf
      for (int v : comp) dist[v] = -1;
      t.emit_only = true;
      stack.push_back(std::move(t));
      continue;
    }
    // middle level by cumulative size
    std::vector<int> count(depth + 1, 0);
    for (int v : lev) count[dist[v]]++;
    int j = 1, acc = count[0];
    const int half = (int)lev.size() / 2;
    while (j < depth - 1 && acc + count[j] < half) acc += count[j++];
    const int aid = next_region++, bid = next_region++, sid = next_region++;
    Task A, B, S;
    A.id = aid; B.id = bid; S.id = sid;
    A.emit_only = B.emit_only = false;
    S.emit_only = true;
    for (int v : lev) {
      const int d = dist[v];
      if (d < j) A.verts.push_back(v);
      else if (d > j) B.verts.push_back(v);
      else {
        bool touches_next = false;
        for (int w : adj[v])
          if (region[w] == t.id && dist[w] == j + 1) { touches_next = true; break; }
        (touches_next ? S.verts : A.verts).push_back(v);
      }
    }
    for (int v : A.verts) region[v] = aid;
    for (int v : B.verts) region[v] = bid;
    for (int v : S.verts) region[v] = sid;
    for (int v : lev) dist[v] = -1;
    // LIFO: halves are processed (and numbered) before the separator
    stack.push_back(std::move(S));
    if (!B.verts.empty()) stack.push_back(std::move(B));
    if (!A.verts.empty()) stack.push_back(std::move(A));
  }
}

}  // namespace detail

struct PlanOptions {
  int fmax = 64;        // largest static subtree front (own columns + contribution rows)
  int dmax = 48;        // delayed-pivot capacity per front / per root
  int sbuf = 96;        // rows of the shared-memory front buffer
  int relax_zeros = 24; // explicit zeros tolerated when merging a child supernode into its parent
  int merge_max = 16;   // largest front produced by a merge that introduces explicit zeros
  int dslot = 16;       // delayed columns one front may hand to its parent
  int tiny = 16;        // largest static front handled by a single warp
  int medium = 24;      // largest static front handled by a two-warp group
  int tiny_max_children = 8;  // fronts with more children are assembled by the whole CTA (staged fetch)
  int nd_leaf = 48;     // nested dissection stops at components of this many vertices
  int int_leaf = 16;    // interior dissection: pieces of at most this many (paired) columns -- one medium front each
  int int_max = 4096;   // ... attempted only on interiors up to this size (the staged ordering is slower)
  int ordering = 0;     // 0 = cheapest schedule of the three, 1 = minimum degree, 2 = nested dissection, 3 = interior dissection
  int root_delay_max = 1024;  // cap on the root's delayed-pivot slots
  bool pair_weak = true; // order zero-diagonal columns together with a partner (2x2 pivot pre-selection)
  int min_sparse_n = 192;   // blocks smaller than this are kept as one dense front
  double max_density = 0.20; // ... as are blocks whose factor would fill more than this share of n^2/2
};

// rows/cols: lower-triangular positions (row >= col) of the block's input entries inside the front
// [K | border]: row < n is a K entry, row >= n a border entry (row - n = border row index).
// entries with keep[k] == 0 are ignored.  src[k] = relative value index of entry k.
inline PatternPlan build_plan_with(int n, int m, const std::vector<int> &rows, const std::vector<int> &cols,
                                   const std::vector<int> &src, const PlanOptions &opt, bool force_dense,
                                   int dissect, const std::vector<double> *hint = nullptr) {
  PatternPlan P;
  P.n = n;
  P.m = m;
  const size_t ne = rows.size();
  std::vector<char> hold(n, 0);
  std::vector<std::vector<int>> adj(n);
  for (size_t k = 0; k < ne; ++k) {
    const int r = rows[k], c = cols[k];
    if (r >= n) hold[c] = 1;  // border-touched column: must stay in the root
    else if (r != c) { adj[r].push_back(c); adj[c].push_back(r); }
  }
  for (auto &a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  bool dense = force_dense || n < opt.min_sparse_n;
  std::vector<int> order;
  std::vector<std::vector<int>> lstruct;
  if (!dense) {
    // ---- 2x2 pivot pre-selection for KKT structure ----
    // A column whose diagonal is (numerically) zero -- constraint multipliers -- cannot be eliminated on its
    // own: as a leaf of the tree it would be delayed to its parent.  Such a column is paired with a neighbour
    // (the largest coupling entry, preferring columns with a usable diagonal) and the pair is ordered as ONE
    // vertex, so both land in the same front and form a 2x2 pivot there (compressed-graph ordering in the
    // manner of Duff & Pralet).  `hint` carries representative numeric values; without it only columns with
    // no stored diagonal count as weak.
    std::vector<double> dmag(n, 0.0), rmax(n, 0.0);
    std::vector<char> has_diag(n, 0);
    std::vector<std::vector<std::pair<int, double>>> inc(n);
    for (size_t k = 0; k < ne; ++k) {
      const int r = rows[k], c = cols[k];
      if (r >= n) continue;
      const double v = hint ? (*hint)[k] : 1.0;
      if (r == c) { dmag[r] += v; has_diag[r] = 1; }
      else { inc[r].push_back({c, fabs(v)}); inc[c].push_back({r, fabs(v)}); rmax[r] = std::max(rmax[r], fabs(v)); rmax[c] = std::max(rmax[c], fabs(v)); }
    }
    std::vector<char> weak(n, 0);
    int nweak = 0;
    for (int v = 0; v < n; ++v) {
      weak[v] = !has_diag[v] || (hint && fabs(dmag[v]) <= 1e-8 * rmax[v]);
      nweak += weak[v];
    }
    std::vector<int> mate(n, -1);
    if (opt.pair_weak && nweak > 0) {
      std::vector<int> wl;
      for (int v = 0; v < n; ++v)
        if (weak[v]) wl.push_back(v);
      std::stable_sort(wl.begin(), wl.end(), [&](int a, int b) { return inc[a].size() < inc[b].size(); });
      for (int v : wl) {
        // border-touched columns stay in the dense root, where Bunch-Kaufman needs no help; pairing them
        // would only drag their partner into the root as well
        if (mate[v] >= 0 || hold[v]) continue;
        int best = -1;
        double bv = 0.0;
        bool bstrong = false;
        for (auto &e : inc[v]) {
          const int w = e.first;
          if (mate[w] >= 0 || w == v || hold[w] || e.second <= 0.0) continue;
          const bool strong = !weak[w];
          if (best < 0 || (strong && !bstrong) || (strong == bstrong && (e.second > bv || (e.second == bv && w < best)))) {
            best = w; bv = e.second; bstrong = strong;
          }
        }
        if (best >= 0) { mate[v] = best; mate[best] = v; }
      }
    }
    // contracted graph: a pair is one vertex
    std::vector<int> cid(n, -1);
    std::vector<std::vector<int>> members;
    for (int v = 0; v < n; ++v) {
      if (cid[v] >= 0) continue;
      cid[v] = (int)members.size();
      if (mate[v] >= 0) {
        cid[mate[v]] = cid[v];
        // the member with the usable diagonal first
        if (weak[v] && !weak[mate[v]]) members.push_back({mate[v], v}); else members.push_back({v, mate[v]});
      } else {
        members.push_back({v});
      }
    }
    const int nc2 = (int)members.size();
    std::vector<std::vector<int>> cadj(nc2);
    std::vector<char> chold(nc2, 0);
    for (int v = 0; v < n; ++v) {
      if (hold[v]) chold[cid[v]] = 1;
      for (int w : adj[v])
        if (cid[w] != cid[v]) cadj[cid[v]].push_back(cid[w]);
    }
    for (auto &a : cadj) {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
    for (int c = 0; c < nc2; ++c)
      if (chold[c])
        for (int v : members[c]) hold[v] = 1;  // a pair with a border-touched column stays in the root whole
    std::vector<int> stage, corder;
    std::vector<std::vector<int>> cstruct;
    if (dissect == 1) {
      detail::nested_dissection_stages(nc2, cadj, chold, opt.nd_leaf, stage);
    } else if (dissect == 2) {
      // Interior dissection.  Minimum degree peels a banded or path-like remainder from its ends, which leaves a
      // chain of fronts -- one per level of the kernels' level loop.  Here the lowest two generations of the
      // elimination tree keep their minimum-degree order, and only the graph of what remains (the interior, with
      // the fill of the eliminated generations as cliques) is dissected, so that the chain becomes a shallow tree
      // of independent pieces joined by small separators.
      std::vector<int> o1;
      std::vector<std::vector<int>> s1;
      detail::minimum_degree(nc2, cadj, chold, std::vector<int>(), o1, s1);
      std::vector<int> pos1(nc2, -1), height(nc2, 0);
      for (size_t k = 0; k < o1.size(); ++k) pos1[o1[k]] = (int)k;
      for (int c : o1) {
        int par = -1;
        for (int w : s1[c])
          if (pos1[w] >= 0 && (par < 0 || pos1[w] < pos1[par])) par = w;
        if (par >= 0) height[par] = std::max(height[par], height[c] + 1);
      }
      std::vector<int> iid(nc2, -1), ivert;
      for (int c : o1)
        if (height[c] >= 2) { iid[c] = (int)ivert.size(); ivert.push_back(c); }
      const int ni = (int)ivert.size();
      stage.assign(nc2, 0);
      if (ni > opt.int_leaf && ni <= opt.int_max) {
        std::vector<std::vector<int>> iadj(ni);
        for (int c = 0; c < nc2; ++c) {
          if (pos1[c] < 0) continue;  // held
          if (iid[c] >= 0) {
            for (int w : cadj[c])
              if (iid[w] >= 0) iadj[iid[c]].push_back(iid[w]);
          } else {
            // fill left by an eliminated lower-generation vertex: a clique on its interior neighbours
            std::vector<int> t;
            for (int w : s1[c])
              if (iid[w] >= 0) t.push_back(iid[w]);
            for (size_t a = 0; a < t.size(); ++a)
              for (size_t b = 0; b < t.size(); ++b)
                if (a != b) iadj[t[a]].push_back(t[b]);
          }
        }
        for (auto &a : iadj) {
          std::sort(a.begin(), a.end());
          a.erase(std::unique(a.begin(), a.end()), a.end());
        }
        std::vector<int> istage;
        detail::nested_dissection_stages(ni, iadj, std::vector<char>(ni, 0), opt.int_leaf, istage);
        for (int k = 0; k < ni; ++k) stage[ivert[k]] = 1 + std::max(istage[k], 0);
      } else {
        stage.clear();  // nothing to dissect: plain minimum degree
      }
    }
    detail::minimum_degree(nc2, cadj, chold, stage, corder, cstruct);
    order.clear();
    lstruct.assign(n, {});
    for (int c : corder) {
      std::vector<int> reach;
      for (int d : cstruct[c])
        for (int v : members[d]) reach.push_back(v);
      std::sort(reach.begin(), reach.end());
      if (members[c].size() == 2) {
        const int a = members[c][0], b = members[c][1];
        order.push_back(a);
        order.push_back(b);
        lstruct[a] = reach;
        lstruct[a].push_back(b);
        std::sort(lstruct[a].begin(), lstruct[a].end());
        lstruct[b] = reach;
      } else {
        order.push_back(members[c][0]);
        lstruct[members[c][0]] = reach;
      }
    }
    double fill = 0;
    for (int p : order) fill += (double)lstruct[p].size() + 1;
    const double held = (double)(n - (int)order.size());
    fill += held * (held + 1) / 2;
    if (fill > opt.max_density * 0.5 * (double)n * n) dense = true;
  }
  if (dense) {
    order.clear();
    lstruct.assign(n, {});
  }
  std::vector<int> pos(n, -1);
  for (size_t k = 0; k < order.size(); ++k) pos[order[k]] = (int)k;
  const int big = 1 << 30;
  auto epos = [&](int v) { return pos[v] < 0 ? big + v : pos[v]; };  // held columns come last

  // ---- supernodes by relaxed amalgamation along the elimination tree ----
  const int ne_cols = (int)order.size();
  std::vector<int> sn_of(n, -1), colparent(n, -1);
  struct SN { std::vector<int> cols, rows; int parent = -1; bool dead = false; bool top = false; };
  std::vector<SN> sn(ne_cols);
  for (int k = 0; k < ne_cols; ++k) {
    const int p = order[k];
    sn[k].cols = {p};
    sn[k].rows = lstruct[p];
    std::sort(sn[k].rows.begin(), sn[k].rows.end(), [&](int a, int b) { return epos(a) < epos(b); });
    sn_of[p] = k;
    colparent[p] = sn[k].rows.empty() ? -1 : sn[k].rows.front();
  }
  auto find = [&](int s) { while (sn[s].dead) s = sn[s].parent; return s; };
  // singleton tree, heights (children precede parents in elimination order)
  std::vector<std::vector<int>> ch(ne_cols);
  std::vector<int> height(ne_cols, 0);
  for (int k = 0; k < ne_cols; ++k) {
    const int q = colparent[sn[k].cols.back()];
    sn[k].parent = (q < 0 || pos[q] < 0) ? -1 : pos[q];
    if (sn[k].parent >= 0) {
      ch[sn[k].parent].push_back(k);
      height[sn[k].parent] = std::max(height[sn[k].parent], height[k] + 1);
    }
  }
  // Greedy amalgamation from the top down: a node repeatedly absorbs its TALLEST child (the one on
  // the critical path), which shortens the level schedule of the numeric kernels; explicit zeros are
  // bounded, and merged fronts stay small enough for one warp unless the merge is free of zeros.
  for (int p2 = ne_cols - 1; p2 >= 0; --p2) {
    if (sn[p2].dead) continue;
    while (!ch[p2].empty()) {
      int bi = 0;
      for (int i = 1; i < (int)ch[p2].size(); ++i) {
        const int a = ch[p2][i], b = ch[p2][bi];
        if (height[a] > height[b] || (height[a] == height[b] && a > b)) bi = i;
      }
      const int c = ch[p2][bi];
      const int nc = (int)sn[c].cols.size(), ncb = (int)sn[c].rows.size();
      const int pc = (int)sn[p2].cols.size(), pcb = (int)sn[p2].rows.size();
      const long zeros = (long)nc * (pc + pcb - ncb);
      const int size = nc + pc + pcb;
      bool ok = size <= opt.fmax;
      if (ok && zeros > 0) {
        const bool small_ok = size <= opt.merge_max && (zeros <= opt.relax_zeros || zeros * 4 <= (long)(nc + pc) * size);
        // an only child (or the only child that is not a leaf) is a link of a chain: no parallelism is lost by merging it, one level of the schedule is
        // saved, and a third of the merged columns' area in explicit zeros costs less than that level
        int inner = 0;  // children that are not leaves of the tree
        for (int g : ch[p2]) inner += height[g] > 0;
        // (not when the merge would push a warp- or group-sized parent into the whole-CTA class: those run many
        // at a time, whole-CTA fronts one after the other)
        const bool chain_ok = (ch[p2].size() == 1 || (inner == 1 && height[c] > 0)) && zeros <= 2L * size &&
                              zeros * 3 <= (long)(nc + pc) * size && (size <= opt.medium || pc + pcb > opt.medium);
        ok = small_ok || chain_ok;
      }
      if (!ok) break;
      std::vector<int> merged(sn[c].cols);
      merged.insert(merged.end(), sn[p2].cols.begin(), sn[p2].cols.end());
      sn[p2].cols.swap(merged);
      sn[c].dead = true;
      sn[c].parent = p2;  // forwarding address
      for (int col : sn[c].cols) sn_of[col] = p2;
      ch[p2].erase(ch[p2].begin() + bi);
      for (int g : ch[c]) { sn[g].parent = p2; ch[p2].push_back(g); }
      ch[c].clear();
    }
  }
  // resolve parents after merging
  for (int k = 0; k < ne_cols; ++k)
    if (!sn[k].dead && sn[k].parent >= 0) sn[k].parent = find(sn[k].parent);
  // ---- top of the tree: big fronts and everything above them go to the dense root ----
  for (int k = 0; k < ne_cols; ++k) {
    if (sn[k].dead) continue;
    if ((int)(sn[k].cols.size() + sn[k].rows.size()) > opt.fmax) sn[k].top = true;
    if (sn[k].top && sn[k].parent >= 0) sn[sn[k].parent].top = true;  // parents come later in elimination order
  }
  // elimination order is topological (children first), so one forward pass closes T upwards
  // ---- DFS postorder of the surviving subtree supernodes ----
  std::vector<std::vector<int>> kids(ne_cols);
  std::vector<int> roots;
  for (int k = 0; k < ne_cols; ++k) {
    if (sn[k].dead || sn[k].top) continue;
    const int p = sn[k].parent;
    if (p >= 0 && !sn[p].top) kids[p].push_back(k); else roots.push_back(k);
  }
  std::vector<int> post, newid(ne_cols, -1);
  {
    std::vector<std::pair<int, size_t>> st;
    for (int r : roots) {
      st.push_back({r, 0});
      while (!st.empty()) {
        auto &[v, i] = st.back();
        if (i < kids[v].size()) { const int c = kids[v][i++]; st.push_back({c, 0}); }
        else { newid[v] = (int)post.size(); post.push_back(v); st.pop_back(); }
      }
    }
  }
  // ---- root columns ----
  std::vector<int> rootpos(n, -1);
  for (int k = 0; k < ne_cols; ++k)
    if (!sn[k].dead && sn[k].top)
      for (int c : sn[k].cols) { rootpos[c] = (int)P.rootcols.size(); P.rootcols.push_back(c); }
  for (int v = 0; v < n; ++v)
    if (pos[v] < 0) { rootpos[v] = (int)P.rootcols.size(); P.rootcols.push_back(v); }
  P.nT = (int)P.rootcols.size();
  P.ns = (int)post.size();
  // delayed columns from ALL children of the root meet here: capacity grows with the block (unused slots cost
  // nothing in the factorisation, the root's pivot count is set on the device)
  P.DR = P.ns > 0 ? std::max(opt.dmax, std::min(opt.root_delay_max, n / 32)) : 0;

  // ---- static per-supernode tables ----
  P.col_ptr.assign(1, 0);
  P.row_ptr.assign(1, 0);
  P.parent.resize(P.ns);
  P.nchild.assign(P.ns, 0);
  P.dcap.resize(P.ns);
  P.l_off.resize(P.ns);
  P.fid_off.resize(P.ns);
  P.fs_off.resize(P.ns);
  std::vector<int> below(P.ns, 0);  // columns strictly below each supernode
  std::vector<int> localpos(n, -1);
  for (int s = 0; s < P.ns; ++s) {
    const SN &S = sn[post[s]];
    const int pp = (S.parent >= 0 && !sn[S.parent].top) ? newid[S.parent] : -1;
    P.parent[s] = pp;
    if (pp >= 0) { P.nchild[pp]++; }
    P.cols.insert(P.cols.end(), S.cols.begin(), S.cols.end());
    P.rows.insert(P.rows.end(), S.rows.begin(), S.rows.end());
    P.col_ptr.push_back((int)P.cols.size());
    P.row_ptr.push_back((int)P.rows.size());
  }
  for (int s = 0; s < P.ns; ++s) {
    const int nc = P.col_ptr[s + 1] - P.col_ptr[s];
    if (P.parent[s] >= 0) below[P.parent[s]] += below[s] + nc;
  }
  P.rel.resize(P.rows.size());
  for (int s = 0; s < P.ns; ++s) {
    const int pp = P.parent[s];
    if (pp >= 0) {
      int k = 0;
      for (int i = P.col_ptr[pp]; i < P.col_ptr[pp + 1]; ++i) localpos[P.cols[i]] = k++;
      for (int i = P.row_ptr[pp]; i < P.row_ptr[pp + 1]; ++i) localpos[P.rows[i]] = k++;
      for (int i = P.row_ptr[s]; i < P.row_ptr[s + 1]; ++i) P.rel[i] = localpos[P.rows[i]];
    } else {
      for (int i = P.row_ptr[s]; i < P.row_ptr[s + 1]; ++i) P.rel[i] = rootpos[P.rows[i]];
    }
  }
  // children lists and levels
  P.child_ptr.assign(P.ns + 1, 0);
  for (int s = 0; s < P.ns; ++s)
    if (P.parent[s] >= 0) P.child_ptr[P.parent[s] + 1]++; else P.root_children.push_back(s);
  for (int s = 0; s < P.ns; ++s) P.child_ptr[s + 1] += P.child_ptr[s];
  P.child_idx.resize(P.child_ptr[P.ns]);
  {
    std::vector<int> fillp(P.child_ptr.begin(), P.child_ptr.end() - 1);
    for (int s = 0; s < P.ns; ++s)
      if (P.parent[s] >= 0) P.child_idx[fillp[P.parent[s]]++] = s;
  }
  std::vector<int> level(P.ns, 0);
  for (int s = 0; s < P.ns; ++s)
    if (P.parent[s] >= 0) level[P.parent[s]] = std::max(level[P.parent[s]], level[s] + 1);
  // capacities and storage layout (children precede parents in postorder)
  P.dslot.resize(P.ns);
  P.cb_off.resize(P.ns);
  P.vec_off.resize(P.ns);
  std::vector<char> cls(P.ns, 2);  // 0 tiny, 1 medium, 2 big
  for (int s = 0; s < P.ns; ++s) {
    const int nc = P.col_ptr[s + 1] - P.col_ptr[s], ncb = P.row_ptr[s + 1] - P.row_ptr[s];
    int incoming = 0;
    for (int k = P.child_ptr[s]; k < P.child_ptr[s + 1]; ++k) incoming += P.dslot[P.child_idx[k]];
    P.dcap[s] = std::max(0, std::min(std::min(opt.dmax, incoming), opt.sbuf - nc - ncb));
    P.dslot[s] = std::min(opt.dslot, nc + P.dcap[s]);
    const int cap_rows = nc + P.dcap[s] + ncb, cap_fs = nc + P.dcap[s];
    P.l_off[s] = P.l_total;
    P.l_total += (int64_t)cap_rows * cap_fs;
    P.fid_off[s] = P.fid_total;
    P.fid_total += cap_rows;
    P.fs_off[s] = P.fs_total;
    P.fs_total += cap_fs;
    const int cbdim = ncb + P.dslot[s];
    P.cb_off[s] = P.cb_total;
    P.cb_total += (int64_t)cbdim * cbdim;
    P.vec_off[s] = P.vec_total;
    P.vec_total += cbdim;
    P.max_front = std::max(P.max_front, cap_rows);
    P.nnz_l += (int64_t)nc * (nc + 1) / 2 + (int64_t)nc * ncb;
    const int nch = P.child_ptr[s + 1] - P.child_ptr[s];
    cls[s] = ((nc + ncb) <= opt.tiny && nch < opt.tiny_max_children) ? 0 : ((nc + ncb) <= opt.medium ? 1 : 2);
    P.nlevels = std::max(P.nlevels, level[s] + 1);
  }
  P.stack_cap = P.cb_total;
  std::vector<int> *ptrs[3] = {&P.tiny_ptr, &P.med_ptr, &P.big_ptr};
  std::vector<int> *idxs[3] = {&P.tiny_idx, &P.med_idx, &P.big_idx};
  for (int c = 0; c < 3; ++c) {
    ptrs[c]->assign(P.nlevels + 1, 0);
    for (int s = 0; s < P.ns; ++s)
      if (cls[s] == c) (*ptrs[c])[level[s] + 1]++;
    for (int l = 0; l < P.nlevels; ++l) (*ptrs[c])[l + 1] += (*ptrs[c])[l];
    idxs[c]->resize((*ptrs[c])[P.nlevels]);
    std::vector<int> fillp(ptrs[c]->begin(), ptrs[c]->end() - 1);
    for (int s = 0; s < P.ns; ++s)
      if (cls[s] == c) (*idxs[c])[fillp[level[s]]++] = s;
  }

  // ---- where every input entry goes ----
  struct Tgt { int s, row, col, src; };
  std::vector<Tgt> tg;
  tg.reserve(ne);
  for (int s = 0; s < P.ns; ++s) {
    for (int i = P.col_ptr[s]; i < P.col_ptr[s + 1]; ++i) sn_of[P.cols[i]] = s;  // reuse: column -> new supernode id
  }
  std::vector<char> in_sub(n, 0);
  for (int c : P.cols) in_sub[c] = 1;
  // local position of a variable inside supernode s: computed on demand via a scratch table
  std::vector<int> lp(n, -1);
  std::vector<std::vector<int>> ent_of(P.ns);
  for (size_t k = 0; k < ne; ++k) {
    const int r = rows[k], c = cols[k];
    if (r >= n) {  // border entry -> root, below the delayed slots
      P.root_row.push_back(P.nT + P.DR + (r - n));
      P.root_col.push_back(rootpos[c]);
      P.root_src.push_back(src[k]);
      continue;
    }
    const bool rs = in_sub[r], cs = in_sub[c];
    if (!rs && !cs) {
      const int a = rootpos[r], b = rootpos[c];
      P.root_row.push_back(std::max(a, b));
      P.root_col.push_back(std::min(a, b));
      P.root_src.push_back(src[k]);
      continue;
    }
    // the entry belongs to the front of whichever endpoint is eliminated first
    const int first = (!cs || (rs && epos(r) < epos(c))) ? r : c;
    ent_of[sn_of[first]].push_back((int)k);
  }
  P.ent_ptr.assign(1, 0);
  P.tgt_src_ptr.assign(1, 0);
  for (int s = 0; s < P.ns; ++s) {
    int kk = 0;
    for (int i = P.col_ptr[s]; i < P.col_ptr[s + 1]; ++i) lp[P.cols[i]] = kk++;
    for (int i = P.row_ptr[s]; i < P.row_ptr[s + 1]; ++i) lp[P.rows[i]] = kk++;
    std::vector<Tgt> loc;
    for (int k : ent_of[s]) {
      int a = lp[rows[k]], b = lp[cols[k]];
      if (a < b) std::swap(a, b);
      loc.push_back({s, a, b, src[k]});
    }
    std::stable_sort(loc.begin(), loc.end(), [](const Tgt &x, const Tgt &y) {
      return x.col != y.col ? x.col < y.col : x.row < y.row;
    });
    for (size_t i = 0; i < loc.size(); ++i) {
      if (i == 0 || loc[i].row != loc[i - 1].row || loc[i].col != loc[i - 1].col) {
        if (i) P.tgt_src_ptr.push_back((int)P.tgt_src.size());
        P.tgt_row.push_back(loc[i].row);
        P.tgt_col.push_back(loc[i].col);
      }
      P.tgt_src.push_back(loc[i].src);
    }
    if (!loc.empty()) P.tgt_src_ptr.push_back((int)P.tgt_src.size());
    P.ent_ptr.push_back((int)P.tgt_row.size());
    for (int i = P.col_ptr[s]; i < P.col_ptr[s + 1]; ++i) lp[P.cols[i]] = -1;
    for (int i = P.row_ptr[s]; i < P.row_ptr[s + 1]; ++i) lp[P.rows[i]] = -1;
  }
  return P;
}

}  // namespace ppb

namespace ppb {

// Estimated time of the level-scheduled numeric kernels, in "front latencies": every level costs a
// fixed front overhead plus the pivots of its widest front (fronts of a level run concurrently).
inline double plan_schedule_cost(const PatternPlan &P) {
  // per level: warps share the tiny fronts 16 at a time, two-warp groups the medium ones 8 at a time,
  // big fronts run one after the other; a front costs a fixed overhead plus its pivots
  double cost = 0;
  auto width = [&](int s) { return (double)(P.col_ptr[s + 1] - P.col_ptr[s]); };
  for (int l = 1; l < P.nlevels; ++l) {
    double t = 0, m = 0, b = 0, tmax = 0, mmax = 0;
    for (int k = P.tiny_ptr[l]; k < P.tiny_ptr[l + 1]; ++k) { t += 6.0 + width(P.tiny_idx[k]); tmax = std::max(tmax, 6.0 + width(P.tiny_idx[k])); }
    for (int k = P.med_ptr[l]; k < P.med_ptr[l + 1]; ++k) { m += 8.0 + width(P.med_idx[k]); mmax = std::max(mmax, 8.0 + width(P.med_idx[k])); }
    for (int k = P.big_ptr[l]; k < P.big_ptr[l + 1]; ++k) b += 10.0 + 2.0 * width(P.big_idx[k]);
    cost += std::max(t / 16.0, tmax) + std::max(m / 8.0, mmax) + b + 2.0;
  }
  return cost + 0.02 * (double)(P.nT + P.DR) * (P.nT + P.DR) / 64.0;  // dense root panel steps
}

// Minimum degree minimises fill, nested dissection the depth of the tree; build both, keep the one
// with the shorter schedule unless it costs much more fill.
inline PatternPlan build_plan(int n, int m, const std::vector<int> &rows, const std::vector<int> &cols,
                              const std::vector<int> &src, const PlanOptions &opt, bool force_dense,
                              const std::vector<double> *hint = nullptr) {
  PatternPlan md = build_plan_with(n, m, rows, cols, src, opt, force_dense, 0, hint);
  if (force_dense || md.ns == 0 || opt.ordering == 1) return md;
  if (opt.ordering == 2) return build_plan_with(n, m, rows, cols, src, opt, force_dense, 1, hint);
  if (opt.ordering == 3) return build_plan_with(n, m, rows, cols, src, opt, force_dense, 2, hint);
  PatternPlan best = md;
  double best_cost = plan_schedule_cost(md);
  auto consider = [&](PatternPlan &&cand, double fill_factor, int root_slack, double gain) {
    if (cand.ns == 0) return;
    const bool fill_ok = (double)cand.nnz_l <= fill_factor * (double)md.nnz_l + 1000.0 && cand.nT <= md.nT + root_slack;
    const double c = plan_schedule_cost(cand);
    if (fill_ok && c < gain * best_cost) { best = std::move(cand); best_cost = c; }
  };
  // a chain of fronts in the upper tree (more than a handful of levels): dissect the interior only
  if (md.nlevels > 6) consider(build_plan_with(n, m, rows, cols, src, opt, force_dense, 2, hint), 1.5, 16, 0.9);
  // nested dissection of the whole graph only pays when minimum degree left a deep tree (its staged ordering costs
  // 5-10x the time of plain minimum degree on 20 000-row blocks, and it is rejected whenever it inflates the root)
  if (md.nlevels > 24) consider(build_plan_with(n, m, rows, cols, src, opt, force_dense, 1, hint), 2.5, 64, 0.8);
  return best;
}

}  // namespace ppb
