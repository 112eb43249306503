"""Fuzz the host analysis of sparse coupling systems (csrc/coupling.hpp) on CPU: random clique structures -> the level-by-level
numpy walk of tests/test_coupling_plan.py must reproduce S^-1 b and the inertia.  `python tools/fuzz_coupling.py [cases] [seed]`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.schur_oracle import dense_inertia  # noqa: E402
from tests.test_coupling_plan import dense_from, solve_by_levels  # noqa: E402  (helpers only)


def cliques_of(rng, kind, m_c):
    cl = []
    if kind == "chain":
        gs = int(rng.integers(1, 6))
        ov = int(rng.integers(1, 3))
        g = max(m_c // gs, ov + 2)
        m_c = g * gs
        cl = [np.arange(t * gs, (t + ov + 1) * gs) for t in range(g - ov)]
    elif kind == "random":       # random subsets: little structure to find
        for _ in range(int(rng.integers(2, m_c // 2))):
            cl.append(np.sort(rng.choice(m_c, size=int(rng.integers(1, 7)), replace=False)))
    elif kind == "star":         # scenario groups that all share a few first-stage variables + private ones
        shared = np.arange(int(rng.integers(1, 5)))
        pos = shared.size
        while pos < m_c:
            k = int(min(m_c - pos, rng.integers(1, 9)))
            cl.append(np.concatenate([shared, np.arange(pos, pos + k)]))
            pos += k
    elif kind == "tree":         # groups on a random tree, a clique per edge
        gs = int(rng.integers(1, 5))
        g = max(m_c // gs, 3)
        m_c = g * gs
        for v in range(1, g):
            p = int(rng.integers(0, v))
            cl.append(np.sort(np.concatenate([np.arange(p * gs, (p + 1) * gs), np.arange(v * gs, (v + 1) * gs)])))
    elif kind == "grid":         # groups on a 2-D grid, a clique per edge: separators grow
        side = max(int(np.sqrt(m_c / 2)), 3)
        gs = 2
        m_c = side * side * gs
        def grp(i, j): return np.arange((i * side + j) * gs, (i * side + j + 1) * gs)
        for i in range(side):
            for j in range(side):
                if i + 1 < side: cl.append(np.sort(np.concatenate([grp(i, j), grp(i + 1, j)])))
                if j + 1 < side: cl.append(np.sort(np.concatenate([grp(i, j), grp(i, j + 1)])))
    elif kind == "islands":      # disconnected chains + variables in no clique at all + repeated and empty cliques
        pos = 0
        while pos < m_c - 12:
            ln = int(rng.integers(2, 8)); gs = int(rng.integers(1, 4))
            if pos + (ln + 1) * gs > m_c - 6: break
            for t in range(ln):
                cl.append(np.arange(pos + t * gs, pos + (t + 2) * gs))
            pos += (ln + 1) * gs
        if cl:
            cl.append(cl[0].copy())
        cl.append(np.zeros(0, dtype=np.int64))
    return [np.asarray(c, dtype=np.int64) for c in cl], m_c


def one(rng, case):
    kind = ("chain", "random", "star", "tree", "grid", "islands")[case % 6]
    cl, m_c = cliques_of(rng, kind, int(rng.integers(24, 220)))
    q = [(i, i) for i in range(m_c)]
    for _ in range(int(rng.integers(0, 6))):          # a few off-diagonal entries of Q
        r, c = sorted(rng.choice(m_c, size=2, replace=False), reverse=True)
        q.append((int(r), int(c)))
    spd = bool(case % 2)
    S = dense_from([c for c in cl if c.size], m_c, q, rng, shift=0.0)
    if spd:
        S = S + np.diag(np.abs(S).sum(axis=1) + 1.0)
    b = rng.standard_normal(m_c)
    stats = []
    x, inertia = solve_by_levels(S, cl, q, b, min_mc=int(rng.choice([8, 16, 48])), stats=stats)
    tag = f"case {case} kind {kind} m_c {m_c} cliques {len(cl)} spd {spd} levels {[(s[0], s[2], s[3]) for s in stats if s[1]]}"
    ref = np.linalg.solve(S, b)
    cond = np.linalg.cond(S)
    assert np.linalg.norm(x - ref) <= 1e-9 * cond * max(np.linalg.norm(ref), 1.0) * 10, tag
    assert tuple(inertia) == dense_inertia(S, "eigvalsh"), tag
    return tag


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    for c in range(cases):
        tag = one(rng, c)
        if c % 10 == 0 or c < 6:
            print(tag, flush=True)
    print("ok:", cases, "cases")
