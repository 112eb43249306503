"""Fuzz the host side of the plugin on CPU (structure analysis, value gather, right-hand-side packing, solution
unpacking, re-analysis on a pattern change) over the numpy stand-in of the C ABI (tests/fake_backend.py): random
block-bordered systems in many input forms -- COO / CSR / CSC leaves, nested block matrices and vectors, absent border
blocks, absent Q, duplicates, explicit zeros, upper-triangle garbage, ragged and 1 x 1 blocks, fresh objects and
in-place updates between factorisations -- against a dense solve.  `python tools/fuzz_host.py [cases] [seed]`."""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.schur_oracle import dense_inertia  # noqa: E402
from parapint_b200 import B200SchurComplementLinearSolver, LinearSolverStatus  # noqa: E402
from parapint_b200.carriers import BlockMatrix, BlockVector  # noqa: E402
from parapint_b200.regularization import RegularizedKKT  # noqa: E402
from tests.fake_backend import FakeBackend  # noqa: E402


def _leaf(rng, dense, form, lower_only=False):
    """`dense` as a sparse leaf in the given form; duplicates and explicit zeros on request."""
    M = np.tril(dense) if lower_only else dense
    c = sp.coo_matrix(M)
    r, cc, d = c.row, c.col, c.data
    if form.get("dups") and d.size:      # a quarter of the entries stored as two halves
        uk = np.unique(rng.integers(0, d.size, size=max(1, d.size // 4)))
        d = d.copy()
        d[uk] *= 0.5
        r, cc, d = np.concatenate([r, r[uk]]), np.concatenate([cc, cc[uk]]), np.concatenate([d, d[uk]])
    if form.get("zeros"):
        n0 = max(1, M.shape[0] // 3)
        zr = rng.integers(0, M.shape[0], size=n0) if M.shape[0] else np.zeros(0, dtype=int)
        zc = rng.integers(0, M.shape[1], size=n0) if M.shape[1] else np.zeros(0, dtype=int)
        if lower_only:
            zr, zc = np.maximum(zr, zc), np.minimum(zr, zc)
        if zr.size and zc.size:
            r, cc, d = np.concatenate([r, zr]), np.concatenate([cc, zc]), np.concatenate([d, np.zeros(zr.size)])
    if form.get("shuffle") and d.size:
        p = rng.permutation(d.size)
        r, cc, d = r[p], cc[p], d[p]
    out = sp.coo_matrix((d, (r, cc)), shape=M.shape)
    kind = form.get("kind", "coo")
    return out if kind == "coo" else out.tocsr() if kind == "csr" else out.tocsc()


def _nested_matrix(rng, dense, form):
    """A square dense block as a nested 2 x 2 (or 3 x 3) BlockMatrix with COO leaves (some sub-blocks absent when zero)."""
    n = dense.shape[0]
    parts = int(rng.integers(2, 4))
    cuts = np.sort(rng.choice(np.arange(1, n), size=min(parts - 1, n - 1), replace=False)) if n > 1 else np.zeros(0, dtype=int)
    off = np.concatenate(([0], cuts, [n])).astype(int)
    nb = off.size - 1
    B = BlockMatrix(nb, nb)
    for i in range(nb):
        B.set_row_size(i, off[i + 1] - off[i])
        B.set_col_size(i, off[i + 1] - off[i])
        for j in range(nb):
            sub = dense[off[i]:off[i + 1], off[j]:off[j + 1]]
            if np.any(sub) or (i == j and rng.random() < 0.5):
                B.set_block(i, j, _leaf(rng, sub, {"kind": "coo", "zeros": form.get("zeros")}))
    return B


def _nested_vector(rng, flat):
    n = flat.size
    if n < 2:
        return flat.copy()
    cuts = np.sort(rng.choice(np.arange(1, n), size=min(int(rng.integers(1, 4)), n - 1), replace=False))
    off = np.concatenate(([0], cuts, [n])).astype(int)
    v = BlockVector(off.size - 1)
    for k in range(off.size - 1):
        seg = flat[off[k]:off[k + 1]].copy()
        v.set_block(k, _nested_vector(rng, seg) if seg.size > 3 and rng.random() < 0.3 else seg)
    return v


def build(rng, dense, sizes, forms, absent_border, q_absent):
    nb = len(sizes)
    off = np.concatenate(([0], np.cumsum(sizes))).astype(int)
    kkt = BlockMatrix(nb, nb)
    for i in range(nb):
        kkt.set_row_size(i, sizes[i])
        kkt.set_col_size(i, sizes[i])
    for i in range(nb - 1):
        K = dense[off[i]:off[i + 1], off[i]:off[i + 1]]
        f = forms[i]
        if f.get("nested") and sizes[i] > 1:
            kkt.set_block(i, i, _nested_matrix(rng, K, f))
        else:
            Kin = K if not f.get("garbage_upper") else np.tril(K) + np.triu(rng.standard_normal(K.shape), 1)
            kkt.set_block(i, i, _leaf(rng, Kin, f, lower_only=f.get("lower_only", False)))
        if i not in absent_border:
            kkt.set_block(nb - 1, i, _leaf(rng, dense[off[nb - 1]:, off[i]:off[i + 1]], forms[i]))
            if rng.random() < 0.5:     # the upper border may be there too; it is not read
                kkt.set_block(i, nb - 1, _leaf(rng, dense[off[i]:off[i + 1], off[nb - 1]:], {"kind": "coo"}))
    if not q_absent:
        kkt.set_block(nb - 1, nb - 1, _leaf(rng, dense[off[nb - 1]:, off[nb - 1]:], forms[nb - 1]))
    return kkt


def one(rng, case, comm=None):
    """``comm``: a Communicator of several ranks -- every rank draws the same system from the same stream and checks its own
    blocks (round-robin ownership; a rank may own no block at all) and the replicated coupling block."""
    nblk = int(rng.integers(1, 6))
    sizes = [int(rng.integers(1, 14)) for _ in range(nblk)]
    m_c = int(rng.integers(1, 7))
    sizes.append(m_c)
    N = sum(sizes)
    off = np.concatenate(([0], np.cumsum(sizes))).astype(int)
    absent_border = {i for i in range(nblk) if rng.random() < 0.2}
    q_absent = rng.random() < 0.2 and sum(sizes[:nblk]) >= m_c   # Q = 0 needs a border of full row rank
    dense = np.zeros((N, N))
    for i in range(nblk):
        n = sizes[i]
        M = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.5)
        dense[off[i]:off[i + 1], off[i]:off[i + 1]] = M + M.T + np.diag(3.0 + np.abs(M).sum(axis=1) + np.abs(M).sum(axis=0))
        if i not in absent_border:
            A = rng.standard_normal((m_c, n)) * (rng.random((m_c, n)) < 0.5)
            if rng.random() < 0.5 and m_c > 1:
                A[rng.integers(0, m_c)] = 0.0          # a coupling row this block does not touch
            dense[off[nblk]:, off[i]:off[i + 1]] = A
            dense[off[i]:off[i + 1], off[nblk]:] = A.T
    if not q_absent:
        Q = rng.standard_normal((m_c, m_c))
        dense[off[nblk]:, off[nblk]:] = -(Q @ Q.T) - np.eye(m_c)      # a saddle point: indefinite, nonsingular
    else:
        # Q = 0: S = -sum A K^-1 A^T must be nonsingular, i.e. the stacked border needs full row rank: every coupling
        # row gets a dominant entry in a column of its own
        cols = [(i, c) for i in range(nblk) for c in range(sizes[i])]
        for r, k in enumerate(rng.choice(len(cols), size=m_c, replace=False)):
            i, c = cols[k]
            absent_border.discard(i)
            dense[off[nblk] + r, off[i] + c] += 4.0 * m_c + r
            dense[off[i] + c, off[nblk] + r] = dense[off[nblk] + r, off[i] + c]
    kinds = ("coo", "csr", "csc")
    forms = [{"kind": kinds[int(rng.integers(0, 3))], "dups": rng.random() < 0.3, "zeros": rng.random() < 0.3,
              "shuffle": rng.random() < 0.5, "nested": rng.random() < 0.3, "lower_only": rng.random() < 0.3,
              "garbage_upper": rng.random() < 0.2} for _ in range(nblk + 1)]
    for f in forms:
        if f["garbage_upper"]:
            f["dups"] = f["nested"] = False
    kkt = build(rng, dense, sizes, forms, absent_border, q_absent)
    tag = f"case {case}: sizes {sizes} absent border {sorted(absent_border)} q_absent {q_absent} forms {[(f['kind'], int(f['nested'])) for f in forms]}"
    # one case in three goes through device-side inertia correction: the matrix travels as base + three diagonal shifts
    # (class per row: 0 none, 1 / 2 inside the blocks, 3 on the coupling rows), and a retry with new shifts on the same
    # evaluation must re-use the values that are there
    classes = None
    if rng.random() < 0.34:
        classes = ({i: rng.integers(0, 3, size=sizes[i]).astype(np.int8) for i in range(nblk)},
                   (rng.integers(0, 2, size=m_c) * 3 * (0 if q_absent else 1)).astype(np.int8))
    kw = {} if comm is None else {"comm": comm}
    if classes is not None:
        kw["regularization_classes"] = classes
    solver = B200SchurComplementLinearSolver(backend=FakeBackend(), **kw)
    assert solver.do_symbolic_factorization(kkt).status == LinearSolverStatus.successful, tag
    mine = list(solver.local_block_indices) + [nblk]
    want_inertia = dense_inertia(dense, "eigvalsh")
    for rep in range(3):
        b = rng.standard_normal(N)
        rhs = BlockVector(nblk + 1)
        for i in range(nblk + 1):
            seg = b[off[i]:off[i + 1]].copy()
            rhs.set_block(i, _nested_vector(rng, seg) if rng.random() < 0.4 else seg)
        rhs_before = rhs.flatten().copy()
        dense_now = dense
        if classes is None:
            assert solver.do_numeric_factorization(kkt).status == LinearSolverStatus.successful, tag
        else:
            cls_all = np.concatenate([classes[0][i] for i in range(nblk)] + [classes[1]]).astype(int)
            for attempt in range(2):          # second attempt: other shifts, same evaluation (same token)
                uploads = solver.backend.value_uploads()
                shifts = (float(rng.uniform(0.0, 0.5)), float(rng.uniform(-0.5, 0.0)), float(rng.uniform(-0.5, 0.0)))
                if attempt == 0 and rng.random() < 0.3:
                    shifts = (0.0, 0.0, 0.0)
                reg = RegularizedKKT(kkt, classes, shifts, token=("fuzz", case, rep))
                assert solver.do_numeric_factorization(reg).status == LinearSolverStatus.successful, tag
                dense_now = dense + np.diag(np.array((0.0,) + shifts)[cls_all])
                assert solver.get_inertia() == dense_inertia(dense_now, "eigvalsh"), (tag, attempt)
            assert solver.backend.value_uploads() == uploads, tag          # the retry uploaded nothing
            want_inertia = dense_inertia(dense_now, "eigvalsh")
        assert solver.get_inertia() == want_inertia, (tag, solver.get_inertia(), want_inertia)
        x = solver.do_back_solve(rhs)
        ref = np.linalg.solve(dense_now, b)
        for i in mine:
            xi = x.get_block(i)
            xi = xi.flatten() if hasattr(xi, "nblocks") else np.asarray(xi)
            assert np.allclose(xi, ref[off[i]:off[i + 1]], rtol=1e-8, atol=1e-9), (tag, rep, i)
        assert np.array_equal(rhs.flatten(), rhs_before), tag                  # rhs untouched
        for i in mine:                                                          # structure of the rhs preserved
            tb, xb = rhs.get_block(i), x.get_block(i)
            assert hasattr(tb, "nblocks") == hasattr(xb, "nblocks"), (tag, i)
            if hasattr(tb, "nblocks"):
                assert list(tb.block_sizes()) == list(xb.block_sizes()), (tag, i)
        # next round: new values -- in place, as fresh objects with the same pattern, or with a changed pattern
        scale = 1.0 + 0.1 * (rep + 1)
        dense = dense * scale
        want_inertia = dense_inertia(dense, "eigvalsh")
        mode = int(rng.integers(0, 3))
        if mode == 0:       # values updated in place in the very same leaf objects
            def rescale(B):
                if hasattr(B, "bshape"):
                    for i in range(B.bshape[0]):
                        for j in range(B.bshape[1]):
                            if B.get_block(i, j) is not None:
                                rescale(B.get_block(i, j))
                else:
                    B.data *= scale
            rescale(kkt)
        elif mode == 1:     # fresh objects from the same recipe: the same pattern for plain leaves, a new one for
            # leaves with random duplicates / explicit zeros / shuffled entries (the solver must notice by itself)
            kkt = build(rng, dense, sizes, forms, absent_border, q_absent)
        else:               # pattern change: one diagonal block gains explicit zeros on its diagonal
            i = int(rng.integers(0, nblk))
            K = dense[off[i]:off[i + 1], off[i]:off[i + 1]]
            c = sp.coo_matrix(K)
            n = K.shape[0]
            kkt = build(rng, dense, sizes, forms, absent_border, q_absent)
            kkt.set_block(i, i, sp.coo_matrix((np.concatenate([c.data, np.zeros(n)]),
                                               (np.concatenate([c.row, np.arange(n)]), np.concatenate([c.col, np.arange(n)]))),
                                              shape=K.shape))
    return tag


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    for c in range(cases):
        tag = one(rng, c)
        if c % 25 == 0:
            print(tag, flush=True)
    print("ok:", cases, "cases")
