"""numpy emulation of the multifrontal scheme described by a symbolic plan (TEST INFRASTRUCTURE).

Walks the plan exactly as ``subtree_factor_kernel`` does -- original entries by target, extend-add of
the children through ``rel``, elimination of the supernode's own columns, contribution to the parent or
to the dense root front -- but with no pivoting, so it is only meaningful on matrices whose pivots are
safe in any order (the tests use diagonally dominant ones).  It validates the symbolic analysis on CPU.
"""
import numpy as np


def emulate(plan, values, n, m, block=False):
    """Returns the dense root front (lower triangle, before its factorisation) and the eliminated pivots.

    ``block``: eliminate the columns of a supernode all at once (``F22 - F21 F11^-1 F12``, which is what any pivot
    order *inside* the front arrives at) -- for matrices whose fronts need 2x2 pivots, e.g. KKT blocks with zero
    diagonals whose multiplier columns the analysis paired with a partner; ``pivots`` then holds the eigenvalues of
    the ``F11`` blocks (their signs add up to the inertia of the eliminated part)."""
    nT, DR, ns = plan["nT"], plan["DR"], plan["ns"]
    nroot = nT + DR + m
    root = np.zeros((nroot, nroot))
    for r, c, s in zip(plan["root_row"], plan["root_col"], plan["root_src"]):
        root[r, c] += values[s]
    for t in range(DR):
        root[nT + t, nT + t] = 1.0
    cb = {}
    pivots = []
    children = {s: [] for s in range(ns)}
    for s in range(ns):
        if plan["parent"][s] >= 0:
            children[plan["parent"][s]].append(s)
    for s in range(ns):
        c0, c1 = plan["col_ptr"][s], plan["col_ptr"][s + 1]
        r0, r1 = plan["row_ptr"][s], plan["row_ptr"][s + 1]
        nc, ncb = c1 - c0, r1 - r0
        S = nc + ncb
        F = np.zeros((S, S))
        for e in range(plan["ent_ptr"][s], plan["ent_ptr"][s + 1]):
            v = sum(values[k] for k in plan["tgt_src"][plan["tgt_src_ptr"][e]:plan["tgt_src_ptr"][e + 1]])
            F[plan["tgt_row"][e], plan["tgt_col"][e]] += v
        assert len(children[s]) == plan["nchild"][s]
        for c in children[s]:
            rel = plan["rel"][plan["row_ptr"][c]:plan["row_ptr"][c + 1]]
            M = cb.pop(c)
            for i in range(len(rel)):
                for j in range(i + 1):
                    a, b = max(rel[i], rel[j]), min(rel[i], rel[j])
                    F[a, b] += M[i, j]
        F = np.tril(F) + np.tril(F, -1).T
        if block:
            ev = np.linalg.eigvalsh(F[:nc, :nc])
            if nc and np.abs(ev).min() <= 1e-10 * max(np.abs(ev).max(), 1e-300):
                # e.g. a multiplier column that found no partner: the kernels delay such a pivot to the parent front
                raise np.linalg.LinAlgError("pivot block of a front is singular: needs delayed pivots")
            pivots.extend(ev)
            F[nc:, nc:] -= F[nc:, :nc] @ np.linalg.solve(F[:nc, :nc], F[:nc, nc:])
        for k in range(0 if block else nc):
            d = F[k, k]
            pivots.append(d)
            l = F[k + 1:, k] / d
            F[k + 1:, k + 1:] -= np.outer(l, F[k + 1:, k])
        M = np.tril(F[nc:, nc:])
        if plan["parent"][s] >= 0:
            cb[s] = M
        else:
            rel = plan["rel"][r0:r1]
            for i in range(ncb):
                for j in range(i + 1):
                    a, b = max(rel[i], rel[j]), min(rel[i], rel[j])
                    root[a, b] += M[i, j]
    assert not cb
    return root, np.asarray(pivots)
