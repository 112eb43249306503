"""CPU: the host symbolic analysis (minimum degree, supernodes, subtree/root split, entry targets)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.kkt_generator import EstimationModel
from parapint_b200 import native
from tests.multifrontal_emulation import emulate


def _lower_entries(K, A):
    """Entries of the front [K | A] as the solver passes them: lower triangle of K, then the border rows."""
    n = K.shape[0]
    Kc = sp.tril(K).tocoo()
    rows, cols, vals = list(Kc.row), list(Kc.col), list(Kc.data)
    Ac = sp.coo_matrix(A)
    nz = np.unique(Ac.row)
    look = {r: i for i, r in enumerate(nz)}
    rows += [n + look[r] for r in Ac.row]
    cols += list(Ac.col)
    vals += list(Ac.data)
    return np.asarray(rows), np.asarray(cols), np.asarray(vals, dtype=float), len(nz)


def _check_plan_invariants(plan, n, fmax=64, sbuf=96, tiny=16, medium=24, tiny_max_children=8, nent=None):
    """What the subtree kernels rely on: every column eliminated once, postorder, child -> parent maps inside the
    parent's front, static fronts within ``fmax`` and fronts with their delayed-pivot slots within the shared-memory
    buffer (``sbuf`` rows), and the per-level work lists (tiny / medium / big) listing every front exactly once, at its
    own level, after all of its children."""
    ns = plan["ns"]
    level = np.zeros(ns, dtype=int)
    for s in range(ns):
        p = plan["parent"][s]
        if p >= 0:
            level[p] = max(level[p], level[s] + 1)
    listed = np.zeros(ns, dtype=int)
    for name, ptr, idx in (("tiny", plan["tiny_ptr"], plan["tiny_idx"]), ("med", plan["med_ptr"], plan["med_idx"]),
                           ("big", plan["big_ptr"], plan["big_idx"])):
        if ns == 0:
            continue
        assert len(ptr) == plan["nlevels"] + 1 and ptr[-1] == len(idx)
        for lv in range(plan["nlevels"]):
            for s in idx[ptr[lv]:ptr[lv + 1]]:
                listed[s] += 1
                assert level[s] == lv, (name, s, level[s], lv)
                size = (plan["col_ptr"][s + 1] - plan["col_ptr"][s]) + (plan["row_ptr"][s + 1] - plan["row_ptr"][s])
                if name == "tiny":
                    assert size <= tiny and plan["nchild"][s] < tiny_max_children
                elif name == "med":
                    assert size <= medium
    assert np.all(listed == 1)
    for s in range(ns):
        nc = plan["col_ptr"][s + 1] - plan["col_ptr"][s]
        ncb = plan["row_ptr"][s + 1] - plan["row_ptr"][s]
        assert nc >= 1 and nc + ncb <= fmax, (s, nc, ncb, fmax)
        assert 0 <= plan["dcap"][s] and nc + ncb + plan["dcap"][s] <= sbuf
        assert 0 <= plan["dslot"][s] <= nc + plan["dcap"][s]
        kids = plan["child_idx"][plan["child_ptr"][s]:plan["child_ptr"][s + 1]]
        assert len(kids) == plan["nchild"][s] and all(plan["parent"][k] == s for k in kids)
        assert plan["dcap"][s] <= sum(plan["dslot"][k] for k in kids)      # slots only for what the children can delay
    if ns:
        assert plan["max_front"] <= sbuf
        assert sorted(plan["root_children"]) == [s for s in range(ns) if plan["parent"][s] < 0]
        # entry targets: inside the lower triangle of their front, in one of the front's own columns
        for s in range(ns):
            nc = plan["col_ptr"][s + 1] - plan["col_ptr"][s]
            ncb = plan["row_ptr"][s + 1] - plan["row_ptr"][s]
            e0, e1 = plan["ent_ptr"][s], plan["ent_ptr"][s + 1]
            r, c = plan["tgt_row"][e0:e1], plan["tgt_col"][e0:e1]
            assert np.all((0 <= c) & (c < nc) & (c <= r) & (r < nc + ncb)), s
        nroot = plan["nT"] + plan["DR"] + plan["m"]
        rr, rc = plan["root_row"], plan["root_col"]
        assert np.all((0 <= rc) & (rc <= rr) & (rr < nroot) & (rc < plan["nT"]))
        assert np.all((rr < plan["nT"]) | (rr >= plan["nT"] + plan["DR"]))      # nothing lands in the delayed-pivot slots
    if nent is not None and ns:
        # every input entry is used exactly once: in a subtree front or in the root
        used = np.bincount(np.concatenate([plan["tgt_src"], plan["root_src"]]), minlength=nent)
        assert used.size == nent and np.all(used == 1)
    seen = np.zeros(n, dtype=int)
    seen[plan["rootcols"]] += 1
    seen[plan["cols"]] += 1
    assert np.all(seen == 1)  # every column is eliminated exactly once: in a subtree front or in the root
    for s in range(ns):
        p = plan["parent"][s]
        assert p == -1 or p > s  # postorder
        rel = plan["rel"][plan["row_ptr"][s]:plan["row_ptr"][s + 1]]
        assert len(set(rel)) == len(rel) and np.all(rel >= 0)
        if p >= 0:
            size_p = (plan["col_ptr"][p + 1] - plan["col_ptr"][p]) + (plan["row_ptr"][p + 1] - plan["row_ptr"][p])
            assert np.all(rel < size_p)
        else:
            assert np.all(rel < plan["nT"])


@pytest.mark.parametrize("ordering", [0, 1, 2, 3])
@pytest.mark.parametrize("seed,n,m,density", [(0, 300, 7, 0.01), (1, 500, 20, 0.006), (2, 400, 0, 0.005)])
def test_plan_reproduces_schur_complement(seed, n, m, density, ordering):
    """Every ordering (0 = cheapest schedule, 1 = minimum degree, 2 = nested dissection, 3 = interior
    dissection) must yield a plan whose emulated elimination reproduces -A K^-1 A^T."""
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal)
    K = (M + M.T).tolil()
    K.setdiag(np.abs(K).sum(axis=1).A1 + 1.0)  # diagonally dominant: any pivot order is safe
    K = K.tocsr()
    A = np.zeros((max(m, 1), n))
    if m:
        for a in range(m):
            A[a, rng.choice(n, size=3, replace=False)] = rng.standard_normal(3)
    else:
        A = np.zeros((0, n))
    rows, cols, vals, mm = _lower_entries(K, A)
    assert mm == m
    plan = native.build_plan(n, m, rows, cols, min_sparse_n=64, ordering=ordering)
    if ordering in (2, 3) and plan["ns"] == 0:
        assert plan["nT"] == n  # a forced dissection may fill past the density cut-off on an expander-like graph: one dense front
        return
    assert plan["ns"] > 0 and plan["nT"] < n
    _check_plan_invariants(plan, n)
    root, pivots = emulate(plan, vals, n, m)
    assert np.all(pivots > 0)
    # finish the root densely and compare the trailing block with -A K^-1 A^T
    nr = plan["nT"] + plan["DR"]
    R = np.tril(root) + np.tril(root, -1).T
    schur = R[nr:, nr:] - R[nr:, :nr] @ np.linalg.solve(R[:nr, :nr], R[:nr, nr:])
    expect = -A @ np.linalg.solve(K.toarray(), A.T)
    assert np.allclose(schur, expect, rtol=1e-9, atol=1e-11)


def test_plan_generator_block():
    """BASELINE config-2 block (n = 2000, nnz 8176): most columns leave the dense root."""
    m = EstimationModel(1, 150, 6, 50)
    K = m.blocks[0].kkt().tocsr()
    A = m.border().toarray()
    rows, cols, vals, mm = _lower_entries(K, A)
    plan = native.build_plan(2000, mm, rows, cols)
    _check_plan_invariants(plan, 2000)
    assert plan["nT"] >= 50 and plan["nT"] < 400  # the 50 border-touched columns stay in the root
    assert plan["max_front"] <= 96
    assert set(np.where(A.any(axis=0))[0]) <= set(plan["rootcols"])


def test_small_or_dense_blocks_stay_dense():
    n = 40
    K = sp.csr_matrix(np.ones((n, n)))
    rows, cols, vals, mm = _lower_entries(K, np.zeros((0, n)))
    plan = native.build_plan(n, 0, rows, cols)
    assert plan["ns"] == 0 and plan["nT"] == n and plan["DR"] == 0
    assert np.array_equal(plan["rootcols"], np.arange(n))
    n = 400
    rng = np.random.default_rng(0)
    D = rng.standard_normal((n, n))
    rows, cols, vals, mm = _lower_entries(sp.csr_matrix(D + D.T), np.zeros((0, n)))
    plan = native.build_plan(n, 0, rows, cols)
    assert plan["ns"] == 0  # fill would exceed the density cut-off


def test_interior_dissection_shortens_a_chain():
    """A banded block (pentadiagonal, the shape the q-columns of the generator family leave behind): minimum
    degree peels it from the ends into a chain of fronts; the interior dissection must give a shallower level
    schedule, no more root columns, and the same Schur complement."""
    n, m = 600, 4
    rng = np.random.default_rng(5)
    K = sp.diags([np.full(n - 2, 0.3), np.full(n - 1, -0.7), np.full(n, 4.0), np.full(n - 1, -0.7), np.full(n - 2, 0.3)],
                 [-2, -1, 0, 1, 2]).tocsr()
    A = np.zeros((m, n))
    A[np.arange(m), rng.choice(n, size=m, replace=False)] = 1.0
    rows, cols, vals, mm = _lower_entries(K, A)
    md = native.build_plan(n, m, rows, cols, min_sparse_n=64, ordering=1)
    it = native.build_plan(n, m, rows, cols, min_sparse_n=64, ordering=3)
    auto = native.build_plan(n, m, rows, cols, min_sparse_n=64, ordering=0)
    assert it["nlevels"] < md["nlevels"] and it["nT"] <= md["nT"] + 16
    assert auto["nlevels"] <= md["nlevels"]
    for plan in (md, it, auto):
        _check_plan_invariants(plan, n)
        root, pivots = emulate(plan, vals, n, m)
        nr = plan["nT"] + plan["DR"]
        R = np.tril(root) + np.tril(root, -1).T
        schur = R[nr:, nr:] - R[nr:, :nr] @ np.linalg.solve(R[:nr, :nr], R[:nr, nr:])
        assert np.allclose(schur, -A @ np.linalg.solve(K.toarray(), A.T), rtol=1e-9, atol=1e-11)


def test_plan_fuzz_structured_patterns():
    """A slice of ``tools/fuzz_symbolic.py``: banded, arrow, disconnected, grid, diagonal-only and KKT-shaped patterns
    (multiplier columns without a diagonal entry -> ordered as 2x2 pivots with a partner), duplicated and shuffled
    entries, every ordering and front-size cap; plan invariants, the Schur complement and the determinant must come out
    of the numpy walk of the plan."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_symbolic.py")
    spec = importlib.util.spec_from_file_location("fuzz_symbolic", path)
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(11)
    walked = 0
    with np.errstate(all="ignore"):
        for case in range(35):
            walked += not fuzz.one(rng, case).startswith("skipped")
    assert walked >= 30


@pytest.mark.parametrize("shape,want", [
    ((2, 150, 6, 50), {"ns": 931, "nT": 50, "DR": 62, "nnz_l": 10737, "max_front": 86}),           # BASELINE config 2
    ((2, 2000, 4, 2000), {"ns": 8624, "nT": 2082, "DR": 625, "nnz_l": 113961, "max_front": 84}),   # BASELINE config 5
])
def test_generator_plans_are_the_measured_ones(shape, want):
    """The plan of a generator block as the solver builds it (block + its border rows) has the statistics of the plan
    the bench records were measured with (``profiles/bench_r02_n1*.json`` / ``bench_r02_n8.json``: ``symbolic``): a
    change of the ordering or of the amalgamation shows up here before it shows up as a different step time on the GPU."""
    from parapint_b200 import structure
    m = EstimationModel(*shape)
    st = structure.analyse(m.build_kkt())
    sel = st.dest_front == 0
    n, mloc = int(st.block_n[0]), int(st.border_ptr[1] - st.border_ptr[0])
    plan = native.build_plan(n, mloc, st.dest_row[sel], st.dest_col[sel])
    _check_plan_invariants(plan, n, nent=int(sel.sum()))
    got = {k: plan[k] for k in want}
    assert got == want, got
    assert plan["nlevels"] <= 8


@pytest.mark.parametrize("family,want", [
    ("config4", {"ns": 3679, "nT": 1972, "DR": 631, "nnz_l": 274009, "max_front": 96}),    # 20 200-row scenario, 200 first-stage columns
    ("config3", {"ns": 920, "nT": 232, "DR": 312, "nnz_l": 157021, "max_front": 96}),      # first time block, 50 interface states
])
def test_ipm_shaped_plans_are_the_measured_ones(family, want):
    """Family-P blocks at the shapes of BASELINE configs 4 and 3, analysed WITH the values hint as ``pp_symbolic`` does
    (``pp_plan_set_hint``: multiplier columns, whose stored diagonal is an explicit zero, are ordered as 2x2 pivots with
    a partner): the plans have the statistics recorded by the bench legs on the GPU (``profiles/bench_r02_n8.json``,
    ``other_workloads.*.symbolic``) and satisfy the invariants the kernels rely on."""
    from oracle.kkt_families import dynamic_ipm_system, stochastic_ipm_system
    from parapint_b200 import structure
    if family == "config4":
        kkt, _ = stochastic_ipm_system(7, 128, 10000, 8000, 1000, 200, same_pattern=True, local_blocks=[0])
        st = structure.analyse(kkt, 0, 128)
    else:
        kkt, _ = dynamic_ipm_system(9, 256, 5000, 4800, 100, 50, same_pattern=True, local_blocks=[0])
        st = structure.analyse(kkt, 0, 256)
    assert st.n_local == 1
    hint = np.zeros(st.nvals)
    assert structure.gather_values(kkt, st, hint)
    sel = st.dest_front == 0
    n, mloc = int(st.block_n[0]), int(st.border_ptr[1] - st.border_ptr[0])
    plan = native.build_plan(n, mloc, st.dest_row[sel], st.dest_col[sel], values=hint[sel])
    _check_plan_invariants(plan, n, nent=int(sel.sum()))
    got = {k: plan[k] for k in want}
    assert got == want, got
    if family == "config4":    # the hint matters: without it the multiplier columns are not paired and the root grows
        unhinted = native.build_plan(n, mloc, st.dest_row[sel], st.dest_col[sel])
        assert unhinted["nT"] > plan["nT"] and unhinted["ns"] != plan["ns"]
