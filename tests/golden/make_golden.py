"""Emit golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Executes the reference's own ``SchurComplementLinearSolver`` /
``MPISchurComplementLinearSolver`` / ``ScipyInterface`` and its synthetic KKT
generator from ``/root/reference`` (via ``oracle.reference_loader``; nothing is
copied) and stores inputs + outputs as small ``.npz`` files next to this
script.  ``tests/test_oracle.py`` pins the oracle restatement against them and
the ``-m gpu`` tests pin the CUDA path against the oracle and against them.
"""
import os
import sys

import numpy as np
from scipy.sparse import coo_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_loader  # noqa: E402


def _kat_system(ref, sym, q11):
    """8x8 system of reference test_explicit_schur_complement.py:13-38 (sym=False) /
    test_mpi_explicit_schur_complement.py:46-47 (q11=1).  ``sym=True`` replaces the two
    non-symmetric diagonal blocks by their symmetrisation (an LDL^T solver reads one triangle)."""
    b0 = [[1, 1], [1, 1.5]] if sym else [[1, 1], [0, 1]]
    b2 = [[1.5, 1], [1, 1]] if sym else [[1, 0], [1, 1]]
    A = ref.BlockMatrix(4, 4)
    A.set_block(0, 0, coo_matrix(np.array(b0, dtype=np.double)))
    A.set_block(1, 1, coo_matrix(np.array([[1, 0], [0, 1]], dtype=np.double)))
    A.set_block(2, 2, coo_matrix(np.array(b2, dtype=np.double)))
    A.set_block(3, 3, coo_matrix(np.array([[0, 0], [0, q11]], dtype=np.double)))
    A.set_block(3, 0, coo_matrix(np.array([[0, -1], [0, 0]], dtype=np.double)))
    A.set_block(3, 1, coo_matrix(np.array([[-1, 0], [0, -1]], dtype=np.double)))
    A.set_block(3, 2, coo_matrix(np.array([[0, 0], [-1, 0]], dtype=np.double)))
    rhs = ref.BlockVector(4)
    rhs.set_block(0, np.array([1, 0], dtype=np.double))
    rhs.set_block(1, np.array([0, 0], dtype=np.double))
    rhs.set_block(2, np.array([0, 1], dtype=np.double))
    rhs.set_block(3, np.array([1, 1], dtype=np.double))
    return A, rhs


def _dense_of(A):
    full = A.toarray()
    n = full.shape[0]
    low = full[n - 2:, : n - 2]
    full[: n - 2, n - 2:] = low.T
    return full


def _ref_serial(ref, A, rhs, inertia=True):
    n = A.bshape[0] - 1
    s = ref.SchurComplementLinearSolver({i: ref.ScipyInterface(compute_inertia=inertia) for i in range(n)},
                                        ref.ScipyInterface(compute_inertia=inertia))
    st1 = s.do_symbolic_factorization(A)
    st2 = s.do_numeric_factorization(A)
    x = s.do_back_solve(rhs.copy())
    return x, (s.get_inertia() if inertia else None), (st1.status.value, st2.status.value)


def main():
    ref = reference_loader.load()
    out = {}

    # --- known-answer 8x8 systems -------------------------------------------------
    for name, sym, q11 in (("orig", False, 0.0), ("sym", True, 0.0), ("sym_q", True, 1.0), ("orig_q", False, 1.0)):
        A, rhs = _kat_system(ref, sym, q11)
        dense = _dense_of(A)
        x, inertia, status = _ref_serial(ref, A, rhs)
        out[f"kat_{name}_dense"] = dense
        out[f"kat_{name}_rhs"] = rhs.flatten()
        out[f"kat_{name}_x"] = x.flatten()
        out[f"kat_{name}_inertia"] = np.asarray(inertia, dtype=np.int64)
        out[f"kat_{name}_status"] = np.asarray(status, dtype=np.int64)
        assert np.allclose(np.linalg.solve(dense, rhs.flatten()), x.flatten())
    # MPI solver (1-rank communicator): test_mpi_explicit_schur_complement.py:95-115 incl. refactor reuse
    A, rhs = _kat_system(ref, True, 1.0)
    Ampi = ref.MPIBlockMatrix(4, 4, -np.ones((4, 4), dtype=np.int64), ref.comm)
    rmpi = ref.MPIBlockVector(4, -np.ones(4, dtype=np.int64), ref.comm)
    for i in range(4):
        for j in range(4):
            if A.get_block(i, j) is not None:
                Ampi.set_block(i, j, A.get_block(i, j))
        rmpi.set_block(i, rhs.get_block(i))
    s = ref.MPISchurComplementLinearSolver({i: ref.ScipyInterface(compute_inertia=True) for i in range(3)},
                                           ref.ScipyInterface(compute_inertia=True))
    s.do_symbolic_factorization(Ampi)
    s.do_numeric_factorization(Ampi)
    x1 = s.do_back_solve(rmpi).flatten()
    s.do_numeric_factorization(Ampi)
    x2 = s.do_back_solve(rmpi).flatten()
    out["kat_mpi_x"] = x1
    out["kat_mpi_x_refactor"] = x2
    out["kat_mpi_inertia"] = np.asarray(s.get_inertia(), dtype=np.int64)

    # --- 3x3 leaf system: test_linear_solvers.py:13-23,64-79 -----------------------
    mat = coo_matrix(([1, 7, 3, 7, 4, 3, 6.0], ([0, 0, 0, 1, 1, 2, 2], [0, 1, 2, 0, 1, 0, 2])), shape=(3, 3))
    leaf = ref.ScipyInterface(compute_inertia=True)
    zero = mat.copy()
    zero.data.fill(0)
    assert leaf.do_symbolic_factorization(zero).status.value == 0
    assert leaf.do_numeric_factorization(mat).status.value == 0
    out["leaf_dense"] = mat.toarray()
    out["leaf_rhs"] = np.stack([mat * np.array([1, 2, 3.0]), mat * np.array([4, 2, 3.0])])
    out["leaf_x"] = np.stack([leaf.do_back_solve(r) for r in out["leaf_rhs"]])
    out["leaf_inertia"] = np.asarray(leaf.get_inertia(), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **out)

    # --- generator family G ---------------------------------------------------------
    gen = {}
    for tag, args, with_inertia, store_x in (
        ("g_3_20_2_5", (3, 20, 2, 5), True, True),
        ("g_4_60_3_10", (4, 60, 3, 10), True, True),
        ("g_3_500_12_10", (3, 500, 12, 10), False, False),  # reference golden max_err (test_examples.py:76-99)
        ("g_64_150_6_50", (64, 150, 6, 50), False, False),  # BASELINE config 2
    ):
        m = ref.Model(*args, 3)
        kkt, rhs = m.build_kkt(), m.build_rhs()
        x, inertia, status = _ref_serial(ref, kkt, rhs, inertia=with_inertia)
        flat = x.flatten()
        N = args[0]
        k0 = kkt.get_block(0, 0).tocsr()
        k0.sum_duplicates()
        gen[f"{tag}_args"] = np.asarray(args, dtype=np.int64)
        gen[f"{tag}_max_err"] = np.float64(m.check_result(x))
        gen[f"{tag}_xc"] = np.asarray(x.get_block(N)).copy()
        gen[f"{tag}_xsum"] = np.float64(flat.sum())
        gen[f"{tag}_xnorm"] = np.float64(np.linalg.norm(flat))
        gen[f"{tag}_xhead"] = np.stack([np.asarray(x.get_block(i))[:32] for i in range(N)])
        gen[f"{tag}_k0_data_sum"] = np.float64(np.abs(k0.data).sum())
        gen[f"{tag}_k0_nnz"] = np.int64(k0.nnz)
        gen[f"{tag}_rhs_sum"] = np.float64(rhs.flatten().sum())
        if with_inertia:
            gen[f"{tag}_inertia"] = np.asarray(inertia, dtype=np.int64)
        if store_x:
            gen[f"{tag}_x"] = flat
            gen[f"{tag}_kkt_lower"] = np.tril(_full_dense(kkt))
            gen[f"{tag}_rhs"] = rhs.flatten()
        print(tag, "max_err", repr(gen[f"{tag}_max_err"]), "inertia", inertia, "status", status)
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **gen)
    print("fixtures written to", HERE)


def _full_dense(kkt):
    N = kkt.bshape[0] - 1
    full = kkt.toarray()
    sizes = [kkt.get_block(i, i).shape[0] for i in range(N + 1)]
    off = np.concatenate(([0], np.cumsum(sizes)))
    full[: off[N], off[N]:] = full[off[N]:, : off[N]].T
    return full


if __name__ == "__main__":
    main()
