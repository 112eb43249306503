"""Fuzz the host symbolic analysis (csrc/symbolic.hpp) on CPU: random structured patterns -> plan invariants and a numpy
walk of the plan (tests/multifrontal_emulation.py) that must reproduce -A K^-1 A^T.  `python tools/fuzz_symbolic.py [cases] [seed]`."""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from parapint_b200 import native  # noqa: E402
from tests.multifrontal_emulation import emulate  # noqa: E402
from tests.test_symbolic_plan import _check_plan_invariants, _lower_entries  # noqa: E402  (helpers only)


def pattern(rng, kind, n, case=0):
    if kind == "random":
        M = sp.random(n, n, density=rng.uniform(0.5, 6) / n, random_state=rng, data_rvs=rng.standard_normal)
    elif kind == "band":
        w = int(rng.integers(1, 6))
        M = sp.diags([rng.standard_normal(n - k) for k in range(1, w + 1)], list(range(1, w + 1)), shape=(n, n))
    elif kind == "arrow":
        M = sp.lil_matrix((n, n))
        hubs = rng.choice(n, size=int(rng.integers(1, 4)), replace=False)
        for h in hubs:
            idx = rng.choice(n, size=int(rng.integers(n // 4, n // 2)), replace=False)
            M[h, idx] = rng.standard_normal(idx.size)
    elif kind == "blocks":     # disconnected components of very different sizes, some columns with no off-diagonal at all
        M = sp.lil_matrix((n, n))
        pos = 0
        while pos < n:
            b = int(min(n - pos, rng.integers(1, 40)))
            if b > 1 and rng.random() < 0.8:
                S = sp.random(b, b, density=min(1.0, 3.0 / b), random_state=rng, data_rvs=rng.standard_normal)
                M[pos:pos + b, pos:pos + b] = S
            pos += b
    elif kind == "grid":
        g = int(np.sqrt(n))
        n = g * g
        T = sp.diags([np.ones(g - 1)], [1], shape=(g, g))
        M = sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g))
        M = sp.coo_matrix(M)
        M.data = rng.standard_normal(M.nnz)
    elif kind == "diag":
        M = sp.coo_matrix((n, n))
    elif kind == "kkt":
        # [[H, J^T], [J, 0]] with NO diagonal entries in the multiplier columns: they are ordered as 2x2 pivots with a
        # partner (symbolic.hpp, pair_weak); H diagonally dominant, so the paired order is safe without pivoting
        nx = int(n * rng.uniform(0.55, 0.8))
        ne = n - nx
        H = sp.random(nx, nx, density=rng.uniform(0.5, 3) / nx, random_state=rng, data_rvs=rng.standard_normal)
        H = (H + H.T).tolil()
        H.setdiag(np.abs(H).sum(axis=1).A1 + 1.0)
        J = sp.lil_matrix((ne, nx))
        cols = rng.permutation(nx)
        for r in range(ne):
            J[r, cols[r]] = rng.uniform(1.0, 2.0) * rng.choice([-1, 1])      # a matching: full row rank
            # further entries only in variables outside the matching (or, in one case out of four, anywhere: then
            # some multiplier column may find no free partner and has to be delayed, which the walk skips)
            pool = cols[ne:] if (case // 7) % 4 and nx > ne else cols
            for c in rng.choice(pool, size=min(pool.size, int(rng.integers(0, 3))), replace=False):
                if c != cols[r]:
                    J[r, c] = 0.3 * rng.standard_normal()
        return sp.bmat([[H, J.T], [J, None]]).tocsr(), n      # (the caller may add explicit zero diagonals + a values hint)
    else:
        raise ValueError(kind)
    M = sp.csr_matrix(M)
    n = M.shape[0]
    K = (M + M.T).tolil()
    K.setdiag(np.abs(K).sum(axis=1).A1 + 1.0)
    return K.tocsr(), n


def one(rng, case):
    kind = ("random", "band", "arrow", "blocks", "grid", "diag", "kkt")[case % 7]
    n = int(rng.integers(70, 700))
    K, n = pattern(rng, kind, n, case)
    m = int(rng.integers(0, 30)) if rng.random() < 0.8 else 0
    A = np.zeros((m, n))
    for a in range(m):
        k = int(rng.integers(1, 5))
        A[a, rng.choice(n, size=k, replace=False)] = rng.standard_normal(k)
    rows, cols, vals, mm = _lower_entries(K, A)
    if mm != m:     # (cannot happen: every row of A has an entry)
        raise AssertionError("border rows")
    if rng.random() < 0.3:   # duplicates and shuffled entry order
        extra = rng.integers(0, rows.size, size=rows.size // 5)
        rows, cols, vals = np.concatenate([rows, rows[extra]]), np.concatenate([cols, cols[extra]]), np.concatenate([vals, np.zeros(extra.size)])
        p = rng.permutation(rows.size)
        rows, cols, vals = rows[p], cols[p], vals[p]
    ordering = int(rng.integers(0, 4))
    fmax = int(rng.choice([-1, 32, 64]))
    hint = None
    if kind == "kkt" and rng.random() < 0.5:
        # the form an interior-point interface hands over: multiplier columns WITH a stored diagonal that is an explicit
        # zero -- only the values hint tells the analysis that they are weak (pp_plan_set_hint / pp_symbolic's hint)
        zero_diag = np.flatnonzero(K.diagonal() == 0.0)
        rows, cols, vals = np.concatenate([rows, zero_diag]), np.concatenate([cols, zero_diag]), np.concatenate([vals, np.zeros(zero_diag.size)])
        hint = vals
    elif rng.random() < 0.2:
        hint = vals            # a hint never changes what a plan computes, only the order
    plan = native.build_plan(n, m, rows, cols, fmax=fmax, min_sparse_n=int(rng.choice([16, 64])), ordering=ordering, values=hint)
    tag = f"case {case} kind {kind} n {n} m {m} ordering {ordering} fmax {fmax} ns {plan['ns']} nT {plan['nT']}"
    if plan["ns"] == 0:
        assert plan["nT"] == n, tag
        return tag
    _check_plan_invariants(plan, n, fmax=64 if fmax < 0 else fmax, nent=rows.size)
    try:
        root, pivots = emulate(plan, vals, n, m, block=(kind == "kkt"))
    except np.linalg.LinAlgError:
        # a multiplier column that found no free partner sits alone in a front: the kernels delay it to the parent,
        # which the numpy walk does not emulate
        return "skipped (delayed pivots needed): " + tag
    assert np.all(pivots != 0) and (kind == "kkt" or np.all(pivots > 0)), tag
    nr = plan["nT"] + plan["DR"]
    R = np.tril(root) + np.tril(root, -1).T
    if m:
        schur = R[nr:, nr:] - R[nr:, :nr] @ np.linalg.solve(R[:nr, :nr], R[:nr, nr:])
        expect = -A @ np.linalg.solve(K.toarray(), A.T)
        assert np.allclose(schur, expect, rtol=1e-6 if kind == "kkt" else 1e-8, atol=1e-8 if kind == "kkt" else 1e-10), tag
    # determinant identity: product of subtree pivots x det(root pivot block) = det(K)
    s1, l1 = np.linalg.slogdet(R[:nr, :nr])
    s2, l2 = np.linalg.slogdet(K.toarray())
    # Haynsworth: inertia of K = signs of the eliminated pivots (eigenvalues of the pivot blocks) + inertia of the root block
    er, ek = np.linalg.eigvalsh(R[:plan["nT"], :plan["nT"]]), np.linalg.eigvalsh(K.toarray())   # (the empty delayed-pivot slots hold 1)
    assert (int((pivots > 0).sum() + (er > 0).sum()), int((pivots < 0).sum() + (er < 0).sum())) == \
        (int((ek > 0).sum()), int((ek < 0).sum())), tag
    s1 *= np.prod(np.sign(pivots))
    assert s1 == s2 and abs(l1 + np.log(np.abs(pivots)).sum() - l2) <= 1e-6 * max(1.0, abs(l2)), tag
    return tag


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    for c in range(cases):
        tag = one(rng, c)
        if c % 20 == 0:
            print(tag, flush=True)
    print("ok:", cases, "cases")
