"""CPU: carriers, symbolic analysis, C-ABI surface and the solver's host logic (fake numpy backend)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.kkt_generator import EstimationModel
from oracle.schur_oracle import SchurOracle, dense_inertia, sym_full
from parapint_b200 import (B200SchurComplementLinearSolver, BlockMatrix, BlockVector, LinearSolverInterface,
                           LinearSolverStatus, native, structure)
from tests.fake_backend import FakeBackend
from tests.helpers import block_vector, bordered_from_dense, random_bordered

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ---------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    """The built library loads (no GPU needed) and exports exactly what include/parapint_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "parapint_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pp_[a-z_0-9]+)\s*\(", body))
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    lib = native.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pp_abi_version() == 1
    assert b"sm_100a" in lib.pp_build_info()


def test_status_codes_match_reference_enum():
    """parapint/linalg/results.py:4-9."""
    assert [s.value for s in LinearSolverStatus] == [0, 1, 2, 3, 4]
    assert [s.name for s in LinearSolverStatus] == ["successful", "not_enough_memory", "singular", "error", "warning"]
    header = open(os.path.join(ROOT, "include", "parapint_b200.h")).read()
    for name, val in (("PP_SUCCESSFUL", 0), ("PP_NOT_ENOUGH_MEMORY", 1), ("PP_SINGULAR", 2), ("PP_ERROR", 3), ("PP_WARNING", 4)):
        assert re.search(rf"{name}\s*=\s*{val}\b", header)


def test_null_handle_calls_fail_cleanly():
    lib = native.load()
    # API misuse is PP_MISUSE (-1), distinct from the run-time status PP_ERROR (3) that ranks must agree on
    header = open(os.path.join(ROOT, "include", "parapint_b200.h")).read()
    assert re.search(r"PP_MISUSE\s*=\s*-1\b", header)
    assert lib.pp_numeric_local(None, None, 0, None, None) == -1
    assert b"symbolic" in lib.pp_last_error()
    assert lib.pp_destroy(None) == 0
    assert lib.pp_set_option(None, b"pivot_tol", 0.0) == -1


def test_ipm_vector_entry_points_reject_misuse_and_need_a_gpu():
    """pp_ipm_* (N3): argument errors are PP_MISUSE before anything is launched; the host module has no CPU path."""
    import torch
    lib = native.load()
    assert lib.pp_ipm_workspace_bytes() >= 8 * 148 * 8 * 8
    assert lib.pp_ipm_fraction_to_boundary(10, 0.99, 0.1, None, None, None, None, None, None, None, None, None) == -1
    assert b"pp_ipm_fraction_to_boundary" in lib.pp_last_error()
    assert lib.pp_ipm_complementarity(-1, 0.1, None, None, None, None, None, None, None, None) == -1
    assert lib.pp_ipm_max_abs(5, None, None, None, None, None) == -1
    assert lib.pp_ipm_axpy(5, None, 7, None, None, None) == -1
    if not torch.cuda.is_available():
        from parapint_b200.ipm_vectors import IpmKernels
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            IpmKernels()


def test_solver_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200SchurComplementLinearSolver()


def test_plugin_surface():
    """base_linear_solver_interface.py:5-56: same method names and call conventions."""
    assert issubclass(B200SchurComplementLinearSolver, LinearSolverInterface)
    import inspect
    for name, params in (("do_symbolic_factorization", ["self", "matrix", "raise_on_error", "timer"]),
                         ("do_numeric_factorization", ["self", "matrix", "raise_on_error", "timer"]),
                         ("do_back_solve", ["self", "rhs", "timer"]), ("get_inertia", ["self"]),
                         ("increase_memory_allocation", ["self", "factor"])):
        assert list(inspect.signature(getattr(B200SchurComplementLinearSolver, name)).parameters) == params


# ---- carriers ------------------------------------------------------------------------------------
def test_block_matrix_nested_and_duplicates():
    inner = BlockMatrix(2, 2)
    inner.set_block(0, 0, sp.coo_matrix(([1.0, 2.0], ([0, 0], [0, 0])), shape=(2, 2)))  # duplicate entry
    inner.set_block(1, 0, sp.coo_matrix(np.array([[0.0, 5.0]])))
    inner.set_col_size(1, 1)
    outer = BlockMatrix(2, 2)
    outer.set_block(0, 0, inner)
    outer.set_block(1, 1, sp.identity(2, format="coo"))
    assert outer.shape == (5, 5) and outer.bshape == (2, 2)
    coo = outer.tocoo()
    assert coo.nnz == 5  # duplicates preserved
    dense = outer.toarray()
    assert dense[0, 0] == 3.0 and dense[2, 1] == 5.0 and dense[3, 3] == 1.0
    assert np.array_equal(outer.transpose().toarray(), dense.T)
    assert outer.copy_structure().get_block(0, 0) is None and outer.copy().toarray()[0, 0] == 3.0


def test_block_vector_structure_ops():
    inner = BlockVector(2)
    inner.set_block(0, np.array([1.0, 2.0]))
    inner.set_block(1, np.array([3.0]))
    v = BlockVector(2)
    v.set_block(0, inner)
    v.set_block(1, np.array([4.0, 5.0]))
    assert v.size == 5 and np.array_equal(v.flatten(), [1, 2, 3, 4, 5])
    w = v - np.ones(5)
    assert w.get_block(0).nblocks == 2 and np.array_equal(w.flatten(), [0, 1, 2, 3, 4])
    c = v.copy_structure()
    c.copyfrom(np.arange(5.0))
    assert np.array_equal(c.get_block(0).get_block(1), [2.0])
    assert np.array_equal(sp.identity(5, format="csr").dot(v), v.flatten())


# ---- symbolic analysis -----------------------------------------------------------------------------
def test_analyse_generator_structure():
    m = EstimationModel(3, 20, 2, 5)
    kkt = m.build_kkt()
    st = structure.analyse(kkt)
    assert st.n_blocks == 3 and st.m_c == 5 and st.local_blocks == [0, 1, 2]
    assert list(st.block_n) == [m.block_dim] * 3
    assert list(st.border_ptr) == [0, 5, 10, 15] and list(st.border_rows[:5]) == [0, 1, 2, 3, 4]
    # border entries land in rows n..n+m-1 of their front, on the last n_theta columns (create_model.py:107-110)
    kind, i, lo, hi = st.segments[1]
    assert kind == "A" and np.array_equal(st.dest_row[lo:hi], m.block_dim + np.arange(5))
    assert np.array_equal(st.dest_col[lo:hi], m.block_dim - 5 + np.arange(5))
    # strict upper triangle of K is dropped
    kind, i, lo, hi = st.segments[0]
    c = kkt.get_block(0, 0).tocoo()
    assert np.array_equal(st.dest_front[lo:hi] >= 0, c.row >= c.col)
    vals = np.zeros(st.nvals)
    assert structure.gather_values(kkt, st, vals)
    assert np.array_equal(vals[lo:hi], c.data)


def test_analyse_partitions_round_robin_and_by_ownership():
    m = EstimationModel(5, 10, 2, 3)
    kkt = m.build_kkt()
    assert structure.analyse(kkt, rank=1, size=2).local_blocks == [1, 3]
    kkt.rank_ownership = -np.ones((6, 6), dtype=np.int64)
    for i, o in enumerate([1, 1, 0, 0, 1]):
        kkt.rank_ownership[i, i] = o
    assert structure.analyse(kkt, rank=1, size=2).local_blocks == [0, 1, 4]
    kkt.rank_ownership[2, 2] = -1
    assert structure.analyse(kkt, rank=0, size=2).local_blocks == [2, 3]  # -1 belongs to rank 0 (mpi...:201-202)


def test_pattern_change_is_detected():
    m = EstimationModel(2, 10, 2, 3)
    kkt = m.build_kkt()
    st = structure.analyse(kkt)
    vals = np.zeros(st.nvals)
    k0 = kkt.get_block(0, 0).tocoo()
    shuffled = sp.coo_matrix((k0.data[::-1], (k0.row[::-1], k0.col[::-1])), shape=k0.shape)
    kkt.set_block(0, 0, shuffled)
    assert not structure.gather_values(kkt, st, vals)


def test_non_square_rejected():
    with pytest.raises(ValueError, match="not square"):
        structure.analyse(BlockMatrix(2, 3))


# ---- solver host logic on the fake backend ---------------------------------------------------------
def _fake_solver(**kw):
    return B200SchurComplementLinearSolver(backend=FakeBackend(), **kw)


def test_host_logic_known_answer(known_answers):
    dense = known_answers["kat_sym_q_dense"]
    kkt = bordered_from_dense(dense, [2, 2, 2, 2])
    rhs = block_vector(known_answers["kat_sym_q_rhs"], [2, 2, 2, 2])
    before = rhs.flatten().copy()
    s = _fake_solver()
    assert s.do_symbolic_factorization(matrix=kkt, raise_on_error=False, timer=None).status == LinearSolverStatus.successful
    assert s.do_numeric_factorization(matrix=kkt, raise_on_error=False, timer=None).status == LinearSolverStatus.successful
    x = s.do_back_solve(rhs)
    assert np.allclose(x.flatten(), known_answers["kat_sym_q_x"])
    assert np.array_equal(rhs.flatten(), before)  # rhs is not mutated (unlike explicit...:141,145)
    assert s.get_inertia() == tuple(known_answers["kat_sym_q_inertia"])
    assert x.nblocks == 4


def test_host_logic_nested_blocks_and_reanalysis():
    """Nested BlockMatrix / BlockVector blocks (sc_ip_interface.py:1248-1266) and a COO re-ordering."""
    rng = np.random.default_rng(5)
    flat = random_bordered(rng, 2, 6, 3)
    kkt = BlockMatrix(3, 3)
    for i in range(2):
        K = flat.get_block(i, i).toarray()
        nest = BlockMatrix(2, 2)
        nest.set_block(0, 0, sp.coo_matrix(K[:4, :4])); nest.set_block(0, 1, sp.coo_matrix(K[:4, 4:]))
        nest.set_block(1, 0, sp.coo_matrix(K[4:, :4])); nest.set_block(1, 1, sp.coo_matrix(K[4:, 4:]))
        kkt.set_block(i, i, nest)
        kkt.set_block(2, i, flat.get_block(2, i))
    kkt.set_block(2, 2, flat.get_block(2, 2))
    rhs = BlockVector(3)
    for i in range(2):
        nest = BlockVector(2)
        nest.set_block(0, rng.standard_normal(4)); nest.set_block(1, rng.standard_normal(2))
        rhs.set_block(i, nest)
    rhs.set_block(2, rng.standard_normal(3))
    s = _fake_solver()
    s.do_symbolic_factorization(kkt)
    s.do_numeric_factorization(kkt)
    x = s.do_back_solve(rhs)
    assert x.get_block(0).nblocks == 2  # structure preserved for set_primal_dual_kkt_solution
    assert np.allclose(sym_full(flat) @ x.flatten(), rhs.flatten())
    # same matrix, different COO order in one block -> silently re-analysed (mumps_interface.py:82-83)
    A = kkt.get_block(2, 0).tocoo()
    kkt.set_block(2, 0, sp.coo_matrix((A.data[::-1], (A.row[::-1], A.col[::-1])), shape=A.shape))
    assert s.do_numeric_factorization(kkt).status == LinearSolverStatus.successful
    assert np.allclose(s.do_back_solve(rhs).flatten(), x.flatten())


def test_host_logic_error_conventions():
    dense = np.zeros((5, 5))
    dense[:2, :2] = [[1.0, 2.0], [2.0, 4.0]]
    dense[2:4, 2:4] = np.eye(2); dense[4, 4] = 1.0; dense[4, 0] = dense[0, 4] = 1.0
    kkt = bordered_from_dense(dense, [2, 2, 1])
    s = _fake_solver()
    with pytest.raises(RuntimeError):
        s.do_numeric_factorization(kkt)  # before symbolic
    s.do_symbolic_factorization(kkt)
    with pytest.raises(RuntimeError):
        s.get_inertia()  # before numeric (ma27_interface.py:197-200)
    assert s.do_numeric_factorization(kkt, raise_on_error=False).status == LinearSolverStatus.singular
    with pytest.raises(RuntimeError, match="singular"):
        s.do_numeric_factorization(kkt, raise_on_error=True)
    with pytest.raises(RuntimeError):
        s.do_back_solve(block_vector(np.ones(5), [2, 2, 1]))
    assert s.increase_memory_allocation(2) is None


def test_host_logic_generator_vs_oracle():
    m = EstimationModel(4, 12, 2, 3)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    s = _fake_solver()
    s.do_symbolic_factorization(kkt)
    s.do_numeric_factorization(kkt)
    o = SchurOracle()
    o.symbolic(kkt); o.numeric(kkt)
    assert np.allclose(s.do_back_solve(rhs).flatten(), o.solve(rhs).flatten(), rtol=1e-9, atol=1e-9)
    assert s.get_inertia() == m.expected_inertia()


def test_host_copier_matches_numpy():
    """pp_host_copy (threaded gather / scatter used to stage the KKT values and the right-hand side) is a pure
    host function: same bytes as numpy slice assignment, pointer table refreshed when the arrays change."""
    from parapint_b200 import native, structure
    from oracle.kkt_generator import EstimationModel
    rng = np.random.default_rng(0)
    arrays = [rng.standard_normal(int(k)) for k in rng.integers(0, 70000, size=40)]
    offs = np.concatenate(([0], np.cumsum([a.size for a in arrays])))
    for threads in (1, 3, 8):
        cp = native.HostCopier(threads)
        out = np.full(offs[-1], np.nan)
        assert cp.copy(arrays, offs[:-1], out)
        assert np.array_equal(out, np.concatenate(arrays))
        back = [np.zeros_like(a) for a in arrays]
        assert cp.copy(back, offs[:-1], out, to_staging=False)
        assert all(np.array_equal(a, b) for a, b in zip(arrays, back))
        arrays2 = [a * 2 for a in arrays]                       # new objects: the cached table must not be reused
        assert cp.copy(arrays2, offs[:-1], out) and np.array_equal(out, 2 * np.concatenate(arrays))
    assert not native.HostCopier(2).copy([np.zeros((4, 4))[:, 0]], [0], np.zeros(4))   # strided: caller falls back
    m = EstimationModel(4, 20, 2, 5)
    kkt, rhs = m.build_kkt(), m.build_rhs()
    st = structure.analyse(kkt, 0, 1)
    a, b = np.zeros(st.nvals), np.zeros(st.nvals)
    cp = native.HostCopier(4)
    for _ in range(2):                                          # second pass takes the validated-leaf fast path
        assert structure.gather_values(kkt, st, a) and structure.gather_values(kkt, st, b, cp)
        assert np.array_equal(a, b)
        kkt.get_block(1, 1).data[:] *= 1.5
    ra, rb = np.zeros(st.local_dim), np.zeros(st.local_dim)
    structure.pack_rhs(rhs, st, ra)
    structure.pack_rhs(rhs, st, rb, native.HostCopier(4))
    assert np.array_equal(ra, rb)


def test_nested_blocks_are_gathered_by_recipe():
    """parapint hands over a freshly built nested 4x4 BlockMatrix per scenario and iteration (interface.py:432-491).
    The values must come out exactly as a flattening with tocoo() gives them, without calling it, and any change of
    the nested pattern (a sub-leaf's indices, an extra sub-block) must be reported so that the symbolic phase is
    repeated."""
    from parapint_b200 import structure
    from parapint_b200.carriers import BlockMatrix
    rng = np.random.default_rng(3)

    def scenario(seed, extra=False, shift=False):
        r = np.random.default_rng(seed)
        H = sp.random(6, 6, density=0.4, random_state=np.random.default_rng(11)).tocoo()
        H = sp.coo_matrix((r.standard_normal(H.nnz), (H.row, H.col)), shape=H.shape)
        J = sp.coo_matrix((r.standard_normal(4), ([0, 1, 2, 2], [0, 2, 3, 5 if not shift else 4])), shape=(3, 6))
        K = BlockMatrix(2, 2)
        K.set_block(0, 0, H)
        K.set_block(1, 0, J)
        K.set_block(0, 1, J.transpose().tocoo())
        if extra:
            K.set_block(1, 1, sp.identity(3, format="coo") * 1e-8)
        else:
            K.set_row_size(1, 3) if hasattr(K, "set_row_size") else None
        return K

    def system(seeds, **kw):
        kkt = BlockMatrix(len(seeds) + 1, len(seeds) + 1)
        for i, sd in enumerate(seeds):
            kkt.set_block(i, i, scenario(sd, **kw))
            kkt.set_block(len(seeds), i, sp.coo_matrix(([-1.0, -1.0], ([0, 1], [0, 1])), shape=(2, 9)))
        kkt.set_block(len(seeds), len(seeds), sp.coo_matrix((2, 2)))
        return kkt

    kkt = system([1, 2, 3])
    st = structure.analyse(kkt, 0, 1)
    assert all(rec is not None for (kind, *_), rec in zip(st.segments, st.recipes) if kind == "K")
    fresh = system([4, 5, 6])                       # new objects, same pattern, new values
    out = np.zeros(st.nvals)
    assert structure.gather_values(fresh, st, out)
    expect = np.concatenate([fresh.get_block(i, i).tocoo().data if kind == "K" else
                             (fresh.get_block(3, i).tocoo().data if kind == "A" else fresh.get_block(3, 3).tocoo().data)
                             for (kind, i, lo, hi) in st.segments])
    assert np.array_equal(out, expect)
    assert not structure.gather_values(system([4, 5, 6], shift=True), st, out)   # a sub-leaf's column index moved
    assert not structure.gather_values(system([4, 5, 6], extra=True), st, out)   # an extra sub-block appeared


def test_fresh_objects_are_validated_by_the_threaded_comparison():
    """A matrix rebuilt from scratch (new leaf objects, new index arrays) is accepted when its pattern equals the
    analysed one and rejected when a single index differs -- through pp_host_equal + pp_host_copy (host-only)."""
    from oracle.kkt_generator import EstimationModel
    m = EstimationModel(5, 40, 3, 6)
    st = structure.analyse(m.build_kkt())
    copier = native.HostCopier(4)
    out = np.zeros(st.nvals)
    fresh = m.build_kkt()
    assert structure.gather_values(fresh, st, out, copier)
    ref = np.zeros(st.nvals)
    assert structure.gather_values(fresh, st, ref)                   # numpy path
    assert np.array_equal(out, ref)
    other = m.build_kkt()
    K = other.get_block(2, 2).tocoo()
    row = K.row.copy()
    row[7], row[8] = row[8], row[7]                                   # same entries, different order: a pattern change
    import scipy.sparse as sp
    other.set_block(2, 2, sp.coo_matrix((K.data, (row, K.col)), shape=K.shape))
    assert not structure.gather_values(other, st, out, copier)
    assert not structure.gather_values(other, st, out)


def test_fastptr_tables():
    """csrc/fastptr.c (CPython helper of the host binding): addresses / byte lengths of many buffers in one call,
    -1 for anything that is not a contiguous buffer of the requested type."""
    from parapint_b200 import _pp_fastptr as fp
    rng = np.random.default_rng(5)
    arrays = [rng.standard_normal(int(k)) for k in rng.integers(0, 500, size=50)]
    ptr, ln = np.zeros(50, dtype=np.uintp), np.zeros(50, dtype=np.int64)
    assert fp.fill(arrays, ptr, ln, "d") == 50
    assert all(int(p) == a.ctypes.data for p, a in zip(ptr, arrays) if a.size) and np.array_equal(ln, [a.nbytes for a in arrays])
    assert fp.fill(tuple(arrays), ptr, ln) == 50                                   # any sequence, any type
    assert fp.fill(arrays[:3] + [np.zeros(4, dtype=np.int64)], ptr, ln, "d") == -1  # wrong type
    assert fp.fill([np.zeros((4, 4))[:, 1]], ptr, ln, "d") == -1                    # strided
    assert fp.fill([object()], ptr, ln) == -1                                       # no buffer
    with pytest.raises(ValueError):
        fp.fill(arrays, np.zeros(3, dtype=np.uintp), ln)
    ro = np.arange(5.0)
    ro.setflags(write=False)
    assert fp.fill([ro], ptr, ln, "d") == 1 and int(ptr[0]) == ro.ctypes.data        # read-only sources are fine
    a = [np.arange(7, dtype=np.int32), np.arange(3, dtype=np.int64)]
    b = [x.copy() for x in a]
    pa, pb = np.zeros(2, dtype=np.uintp), np.zeros(2, dtype=np.uintp)
    assert fp.pairs(a, b, pa, pb, ln) == 2 and list(ln[:2]) == [28, 24]
    assert [int(v) for v in pb] == [x.ctypes.data for x in b]
    assert fp.pairs(a, [b[0], b[1].astype(np.float64)], pa, pb, ln) == -1          # same size, other type
    assert fp.pairs(a, [b[0], np.arange(4, dtype=np.int64)], pa, pb, ln) == -1     # other length
    assert fp.pairs(a, b[:1], pa, pb, ln) == -1
    assert fp.same(a, list(a)) and fp.same(a, tuple(a)) and not fp.same(a, b) and not fp.same(a, a[:1])


def test_gather_paths_agree_with_and_without_the_pointer_helper(monkeypatch):
    """The interpreter builds the same pointer tables as csrc/fastptr.c: flat and nested systems, fresh and re-used
    objects, pattern changes detected on both paths."""
    from oracle.kkt_generator import EstimationModel
    m = EstimationModel(5, 40, 3, 6)
    st = structure.analyse(m.build_kkt())
    fresh, other = m.build_kkt(), m.build_kkt()
    K = other.get_block(2, 2).tocoo()
    col = K.col.copy()
    col[3] += 1 if col[3] + 1 < K.shape[1] else -1
    other.set_block(2, 2, sp.coo_matrix((K.data, (K.row, col)), shape=K.shape))
    ref = np.zeros(st.nvals)
    assert structure.gather_values(fresh, st, ref)
    for helper in (True, False):
        if not helper:
            monkeypatch.setattr(native, "_fp", None)
            monkeypatch.setattr(structure, "_fp", None)
        copier = native.HostCopier(3)
        st.__dict__.pop("_leaf_seen", None)
        for _ in range(2):       # second pass: the objects validated by the first one
            out = np.zeros(st.nvals)
            assert structure.gather_values(fresh, st, out, copier) and np.array_equal(out, ref)
        assert not structure.gather_values(other, st, out, copier)
        assert structure.gather_values(fresh, st, out, copier) and np.array_equal(out, ref)


def test_validated_objects_are_revalidated_when_their_index_tuple_changes():
    """A leaf that passed the comparison is taken on trust afterwards only while it is the same object with the same
    index tuple: a new index tuple on the same object (or a new sub-leaf inside the same nested block) is compared
    again -- and rejected when it differs."""
    from parapint_b200.carriers import BlockMatrix

    def nested(seed):
        r = np.random.default_rng(seed)
        H = sp.coo_matrix((r.standard_normal(5), ([0, 1, 2, 3, 3], [0, 1, 2, 3, 0])), shape=(4, 4))
        J = sp.coo_matrix((r.standard_normal(3), ([0, 1, 1], [0, 2, 3])), shape=(2, 4))
        K = BlockMatrix(2, 2)
        K.set_block(0, 0, H)
        K.set_block(1, 0, J)
        K.set_block(0, 1, J.transpose().tocoo())
        K.set_block(1, 1, sp.coo_matrix((2, 2)))
        return K

    def system(seed):
        kkt = BlockMatrix(3, 3)
        for i in range(2):
            kkt.set_block(i, i, nested(10 * seed + i))
            kkt.set_block(2, i, sp.coo_matrix(([-1.0], ([0], [i])), shape=(1, 6)))
        kkt.set_block(2, 2, sp.coo_matrix((1, 1)))
        return kkt

    st = structure.analyse(system(1))
    copier = native.HostCopier(2)
    kkt = system(2)
    out = np.zeros(st.nvals)
    for scale in (1.0, 3.0):                    # the second call takes the validated objects on trust
        assert structure.gather_values(kkt, st, out, copier)
        assert np.array_equal(out, np.concatenate([kkt.get_block(i, i).tocoo().data if kind == "K" else
                                                   kkt.get_block(2, i).tocoo().data for kind, i, _, _ in st.segments]))
        kkt.get_block(0, 0).get_block(0, 0).data[:] *= scale
    seen = st.__dict__["_leaf_seen"]
    assert all(s is not None for s in seen)
    # same nested object, one sub-leaf replaced by one with another pattern
    J = kkt.get_block(1, 1).get_block(1, 0)
    kkt.get_block(1, 1).set_block(1, 0, sp.coo_matrix((J.data, (J.row, np.array([0, 1, 3]))), shape=J.shape))
    assert not structure.gather_values(kkt, st, out, copier)
    # same flat leaf object, index tuple swapped for an equal one: compared again and accepted
    A = kkt.get_block(2, 0)
    kkt2 = system(2)
    assert structure.gather_values(kkt2, st, out, copier)
    A2 = kkt2.get_block(2, 0)
    if hasattr(A2, "coords"):
        A2.coords = (A2.coords[0].copy(), A2.coords[1].copy())
        assert structure.gather_values(kkt2, st, out, copier)
        A2.coords = (A2.coords[0].copy(), np.array([1], dtype=A2.coords[1].dtype))
        assert not structure.gather_values(kkt2, st, out, copier)


def test_large_solutions_are_unpacked_by_the_copy_pool(monkeypatch):
    """unpack_solution hands a solution of several MB to the threaded copy (pp_host_copy, staging -> array): same
    blocks as the single-threaded path, nothing aliases the staging buffer."""
    from parapint_b200.carriers import BlockVector
    rng = np.random.default_rng(8)
    sizes = [70001, 3, 90000, 50000]
    rhs = BlockVector(len(sizes) + 1)
    for i, n in enumerate(sizes):
        rhs.set_block(i, np.zeros(n))
    inner = BlockVector(2)
    inner.set_block(0, np.zeros(2))
    inner.set_block(1, np.zeros(3))
    rhs.set_block(len(sizes), inner)
    st = structure.Structure(n_blocks=len(sizes), m_c=5, local_blocks=list(range(len(sizes))),
                             block_n=np.asarray(sizes, dtype=np.int32), border_ptr=np.zeros(len(sizes) + 1, dtype=np.int64),
                             border_rows=np.zeros(0, dtype=np.int32), dest_front=np.zeros(0, dtype=np.int32),
                             dest_row=np.zeros(0, dtype=np.int32), dest_col=np.zeros(0, dtype=np.int32),
                             rhs_offsets=np.concatenate(([0], np.cumsum(sizes))).astype(np.int64))
    staging = rng.standard_normal(st.local_dim + 7)      # (the pinned buffer may be longer than the solution)
    x_c = rng.standard_normal(5)
    monkeypatch.setattr(structure, "THREADED_UNPACK_BYTES", 1 << 20)
    copier = native.HostCopier(3)
    calls = []
    real = copier.copy
    copier.copy = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
    a = structure.unpack_solution(rhs, st, staging, x_c, copier)
    b = structure.unpack_solution(rhs, st, staging, x_c)
    assert calls, "the threaded path was not taken"
    for i in range(len(sizes)):
        assert np.array_equal(a.get_block(i), b.get_block(i))
        assert np.array_equal(a.get_block(i), staging[st.rhs_offsets[i]:st.rhs_offsets[i + 1]])
        assert not np.shares_memory(a.get_block(i), staging)
    assert np.array_equal(a.get_block(len(sizes)).flatten(), x_c) and a.get_block(len(sizes)).nblocks == 2


@pytest.mark.timeout(120)
def test_copy_pool_wake_calls_interleave_with_jobs():
    """pp_host_wake only changes WHEN the pool's workers are awake: any interleaving of wake-up calls, gathers,
    scatters and comparisons gives the results of the plain calls (and terminates)."""
    import time
    lib = native.load()
    rng = np.random.default_rng(21)
    assert lib.pp_host_wake(-1, 10) < 0 and lib.pp_host_wake(4, -5) < 0        # misuse is reported, not executed
    cp = native.HostCopier(6)
    cp.wake_us = 50.0
    for it in range(400):
        op = rng.integers(0, 5)
        if op == 0:
            cp.wake(float(rng.choice([1, 20, 200, 3000])))
            if rng.integers(0, 4) == 0:
                time.sleep(float(rng.choice([0.0, 2e-5, 5e-4])))                # sometimes no job follows in time
            continue
        n = int(rng.choice([3, 2000, 40000, 300000]))
        k = int(rng.integers(1, 6))
        arrays = [rng.standard_normal(int(rng.integers(0, n + 1))) for _ in range(k)]
        offs = np.concatenate(([0], np.cumsum([a.size for a in arrays])))
        staging = np.full(offs[-1] + 3, np.nan)
        if op in (1, 2):
            assert cp.copy(arrays, offs[:-1], staging)
            assert np.array_equal(staging[: offs[-1]], np.concatenate(arrays)) and np.all(np.isnan(staging[offs[-1]:]))
        elif op == 3:
            staging[: offs[-1]] = rng.standard_normal(offs[-1])
            back = [np.zeros_like(a) for a in arrays]
            assert cp.copy(back, offs[:-1], staging, to_staging=False)
            assert np.array_equal(np.concatenate(back), staging[: offs[-1]])
        else:
            other = [a.copy() for a in arrays]
            assert cp.all_equal(arrays, other) is True
            big = [j for j, a in enumerate(other) if a.size]
            if big:
                other[big[-1]][-1] += 1.0
                assert cp.all_equal(arrays, other) is False


def test_fuzz_input_forms():
    """A slice of ``tools/fuzz_host.py``: random block-bordered systems as COO / CSR / CSC leaves, nested block matrices
    and vectors, absent border blocks, absent Q, duplicates, explicit zeros, garbage in the unread upper triangle; values
    refreshed in place, as fresh objects and with a changed pattern between factorisations -- solution and inertia
    against dense linear algebra, ``rhs`` untouched, structure of ``rhs`` preserved."""
    import importlib.util

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_host.py")
    spec = importlib.util.spec_from_file_location("fuzz_host", path)
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(21)
    for case in range(40):
        fuzz.one(rng, case)


def test_handle_free_entry_points_reject_malformed_descriptions():
    """``pp_plan_create`` / ``pp_cplan_create`` (the analyses, callable without a GPU) answer PP_MISUSE (-1) to entries
    outside the lower triangle, indices out of range, negative sizes and a decreasing clique pointer -- and survive
    random garbage."""
    lib = native.load()

    def plan(n, m, rows, cols):
        rows, cols = np.ascontiguousarray(rows, dtype=np.int32), np.ascontiguousarray(cols, dtype=np.int32)
        out = ctypes.c_void_p()
        code = lib.pp_plan_create(n, m, rows.size, native.np_ptr(rows), native.np_ptr(cols), -1, -1, 16, ctypes.byref(out))
        if code == 0:
            lib.pp_plan_destroy(out)
        return code

    def cplan(m_c, ptr, rows, qr, qc):
        ptr, rows = np.ascontiguousarray(ptr, dtype=np.int64), np.ascontiguousarray(rows, dtype=np.int32)
        qr, qc = np.ascontiguousarray(qr, dtype=np.int32), np.ascontiguousarray(qc, dtype=np.int32)
        out = ctypes.c_void_p()
        code = lib.pp_cplan_create(m_c, ptr.size - 1, native.np_ptr(ptr), native.np_ptr(rows), qr.size, native.np_ptr(qr),
                                   native.np_ptr(qc), 8, 0.9, ctypes.byref(out))
        if code == 0:
            lib.pp_cplan_destroy(out)
        return code

    assert plan(50, 2, [0, 1, 3], [0, 5, 3]) == -1          # upper-triangle entry
    assert plan(50, 2, [0, 60], [0, 5]) == -1               # row beyond n + m
    assert plan(50, 2, [0, -1], [0, 0]) == -1
    assert plan(50, 2, [51], [50]) == -1                    # border entry in a column that does not exist
    assert plan(-5, 0, [], []) == -1
    assert plan(0, 0, [], []) == 0 and plan(1, 0, [0], [0]) == 0 and plan(100, 0, [], []) == 0
    assert cplan(10, [0, 3], [1, 2, 50], [0], [0]) == -1
    assert cplan(10, [0, 3], [1, -2, 5], [0], [0]) == -1
    assert cplan(10, [0, 3, 2], [1, 2, 5], [0], [0]) == -1  # decreasing pointer
    assert cplan(10, [1, 3], [1, 2, 5], [0], [0]) == -1     # pointer not starting at 0
    assert cplan(10, [0, 3], [1, 2, 5], [11], [0]) == -1
    assert cplan(0, [0], [], [], []) == 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        n, m, k = int(rng.integers(0, 120)), int(rng.integers(0, 5)), int(rng.integers(0, 400))
        assert plan(n, m, rng.integers(-3, n + m + 3, size=k), rng.integers(-3, n + 3, size=k)) in (0, -1)
        m_c, nc = int(rng.integers(0, 80)), int(rng.integers(0, 10))
        ptr = np.concatenate(([0], np.cumsum(rng.integers(0, 8, size=nc))))
        nq = int(rng.integers(0, 20))
        assert cplan(m_c, ptr, rng.integers(-2, m_c + 2, size=int(ptr[-1])), rng.integers(-2, m_c + 2, size=nq),
                     rng.integers(-2, m_c + 2, size=nq)) in (0, -1)
