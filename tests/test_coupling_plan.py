"""Host analysis of a SPARSE coupling system (csrc/coupling.hpp through pp_cplan_*), no GPU needed.

The pattern of S must be what the reference builds (mpi_explicit_schur_complement.py:88-125,228-255: union of
nonzero_rows x nonzero_rows over all blocks, plus Q); the block-bordered view of S at every level must reproduce
S^-1 and the inertia of S when walked in numpy exactly as the kernels walk it (partial LDL^T of every block,
Schur complement into the next level's pattern, recursion)."""
import numpy as np
import pytest

from parapint_b200 import native
from oracle.schur_oracle import dense_inertia


def chain_cliques(n_groups, gs, overlap):
    """Cliques of a chain: clique t covers groups t .. t+overlap (each group `gs` variables)."""
    cl = []
    for t in range(n_groups - overlap):
        cl.append(np.arange(t * gs, (t + overlap + 1) * gs))
    return cl, n_groups * gs


def dense_from(cliques, m_c, q_entries, rng, shift=0.0):
    S = np.zeros((m_c, m_c))
    for rows in cliques:
        B = rng.standard_normal((rows.size, rows.size))
        S[np.ix_(rows, rows)] += B + B.T
    for r, c in q_entries:
        v = rng.standard_normal()
        S[r, c] += v
        if r != c:
            S[c, r] += v
    return S + shift * np.eye(m_c)


def pack(cliques):
    ptr = np.concatenate(([0], np.cumsum([c.size for c in cliques]))).astype(np.int64)
    rows = np.concatenate(cliques).astype(np.int32) if cliques else np.zeros(0, dtype=np.int32)
    return ptr, rows


def solve_by_levels(S, cliques, q_entries, b, min_mc, depth=0, stats=None):
    """Returns (x, inertia) of the symmetric S walking the levels of the analysis."""
    m_c = S.shape[0]
    ptr, rows = pack(cliques)
    q = np.asarray(q_entries, dtype=np.int64).reshape(-1, 2)
    plan = native.coupling_plan(m_c, ptr, rows, q[:, 0], q[:, 1], min_mc=min_mc, max_density=0.6)
    if stats is not None:
        stats.append((m_c, plan["sparse"], plan["n_blocks"], plan["m_next"]))
    if not plan["sparse"]:
        return np.linalg.solve(S, b), np.asarray(dense_inertia(S, "eigvalsh"))
    colptr, rowidx = plan["colptr"], plan["rowidx"]
    colidx = np.repeat(np.arange(m_c), np.diff(colptr))
    # the pattern covers every nonzero of S (lower triangle)
    mask = np.zeros_like(S, dtype=bool)
    mask[rowidx, colidx] = True
    assert not np.any(np.tril(S)[~mask] != 0.0)
    assert np.all(rowidx >= colidx) and np.all(mask[np.arange(m_c), np.arange(m_c)])
    vals = S[rowidx, colidx]
    nb, m_next = plan["n_blocks"], plan["m_next"]
    bn, bptr, brows = plan["block_n"], plan["border_ptr"], plan["border_rows"]
    fronts = [np.zeros((int(bn[k] + bptr[k + 1] - bptr[k]),) * 2) for k in range(nb)] + [np.zeros((m_next, m_next))]
    for v, f, r, c in zip(vals, plan["dest_front"], plan["dest_row"], plan["dest_col"]):
        assert f >= 0 and r >= c
        fronts[f][r, c] += v
    fronts = [np.tril(F) + np.tril(F, -1).T for F in fronts]
    perm_l, perm_c = plan["perm_local"], plan["perm_c"]
    assert sorted(np.concatenate([perm_l, perm_c]).tolist()) == list(range(m_c))
    off = np.concatenate(([0], np.cumsum(bn)))
    Sn = fronts[nb].copy()
    bc = b[perm_c].copy()
    inertia = np.zeros(3, dtype=np.int64)
    Ks, As, Rs = [], [], []
    for k in range(nb):
        n = int(bn[k])
        K, A = fronts[k][:n, :n], fronts[k][n:, :n]
        r = brows[bptr[k]:bptr[k + 1]]
        assert np.all(np.diff(r) > 0)
        Ks.append(K); As.append(A); Rs.append(r)
        inertia += np.asarray(dense_inertia(K, "eigvalsh"))
        Sn[np.ix_(r, r)] -= A @ np.linalg.solve(K, A.T)
        bc[r] -= A @ np.linalg.solve(K, b[perm_l[off[k]:off[k + 1]]])
    next_cliques = [np.asarray(r, dtype=np.int64) for r in Rs if len(r)]
    qn = [(int(r), int(c)) for f, r, c in zip(plan["dest_front"], plan["dest_row"], plan["dest_col"]) if f == nb]
    xc, ine = solve_by_levels(Sn, next_cliques, qn, bc, min_mc, depth + 1, stats)
    x = np.zeros(m_c)
    x[perm_c] = xc
    for k in range(nb):
        x[perm_l[off[k]:off[k + 1]]] = np.linalg.solve(Ks[k], b[perm_l[off[k]:off[k + 1]]] - As[k].T @ xc[Rs[k]])
    return x, inertia + ine


@pytest.mark.parametrize("n_groups,gs,overlap,seed", [(40, 3, 1, 0), (25, 5, 1, 1), (64, 2, 1, 2), (30, 4, 2, 3)])
def test_chain_is_reduced_level_by_level(n_groups, gs, overlap, seed):
    rng = np.random.default_rng(seed)
    cliques, m_c = chain_cliques(n_groups, gs, overlap)
    q = [(i, i) for i in range(m_c)]
    S = dense_from(cliques, m_c, q, rng, shift=0.0)
    b = rng.standard_normal(m_c)
    stats = []
    x, inertia = solve_by_levels(S, cliques, q, b, min_mc=4 * gs, stats=stats)
    assert np.linalg.norm(S @ x - b) <= 1e-8 * np.linalg.norm(b) * np.linalg.cond(S)
    assert np.allclose(x, np.linalg.solve(S, b), rtol=1e-6, atol=1e-8)
    assert tuple(inertia) == dense_inertia(S, "eigvalsh")
    levels = [s for s in stats if s[1]]
    assert len(levels) >= 2                                   # recursion happened
    if overlap == 1:
        # every other group is eliminated: the coupling system shrinks by about half per level (cyclic reduction)
        assert all(s[3] <= 0.67 * s[0] + gs for s in levels)


def test_dynamic_structure_pattern():
    """The coupling pattern of parapint's dynamic structure (sc_ip_interface.py:274-357): block t touches its forward
    multipliers and the coupling variables before it; Q couples every coupling variable with one forward multiplier."""
    N, n_s = 9, 3
    nf = n_s * (N - 1)
    cliques = []
    for t in range(N):
        rows = []
        if t < N - 1:
            rows += list(range(n_s * t, n_s * (t + 1)))                   # forward multipliers of block t
        if t > 0:
            rows += list(range(nf + n_s * (t - 1), nf + n_s * t))        # coupling variables before block t
        cliques.append(np.asarray(sorted(rows)))
    m_c = 2 * nf
    q = [(i, i) for i in range(m_c)] + [(nf + i, i) for i in range(nf)]
    rng = np.random.default_rng(5)
    S = dense_from(cliques, m_c, q, rng)
    ptr, rows = pack(cliques)
    qa = np.asarray(q)
    plan = native.coupling_plan(m_c, ptr, rows, qa[:, 0], qa[:, 1], min_mc=8, max_density=0.9)
    assert plan["sparse"]
    # reference pattern: union of clique squares and Q (mpi_explicit_schur_complement.py:88-125)
    ref = np.zeros((m_c, m_c), dtype=bool)
    for c in cliques:
        ref[np.ix_(c, c)] = True
    for r, c in q:
        ref[r, c] = ref[c, r] = True
    colidx = np.repeat(np.arange(m_c), np.diff(plan["colptr"]))
    got = np.zeros_like(ref)
    got[plan["rowidx"], colidx] = True
    assert np.array_equal(got, np.tril(ref))
    b = rng.standard_normal(m_c)
    x, inertia = solve_by_levels(S, cliques, q, b, min_mc=8)
    assert np.allclose(x, np.linalg.solve(S, b), rtol=1e-6, atol=1e-8)
    assert tuple(inertia) == dense_inertia(S, "eigvalsh")


def test_dense_and_small_systems_stay_dense():
    # stochastic structure: every scenario touches every first-stage variable -> one clique = everything
    m_c = 600
    cl = [np.arange(m_c)] * 5
    ptr, rows = pack(cl)
    plan = native.coupling_plan(m_c, ptr, rows, np.zeros(0), np.zeros(0))
    assert not plan["sparse"]
    # a sparse but small system is not worth a level
    cliques, m_c = chain_cliques(20, 3, 1)
    ptr, rows = pack(cliques)
    assert not native.coupling_plan(m_c, ptr, rows, np.zeros(0), np.zeros(0))["sparse"]
    assert native.coupling_plan(m_c, ptr, rows, np.zeros(0), np.zeros(0), min_mc=10, max_density=0.9)["sparse"]


def test_fuzz_clique_structures():
    """A slice of ``tools/fuzz_coupling.py``: chains, random subsets, stars (scenario groups sharing first-stage
    variables), trees, 2-D grids, disconnected islands with repeated / empty cliques and variables in no clique."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_coupling.py")
    spec = importlib.util.spec_from_file_location("fuzz_coupling", path)
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(4)
    for case in range(24):
        fuzz.one(rng, case)
