"""Synthetic parapint-shaped interior-point KKT systems ("family P"): input synthesis for the
parity tests and the benchmark (TEST INFRASTRUCTURE, see ``oracle/__init__.py``; no solver logic).

* ``stochastic_ipm_system``: two-stage stochastic layout (BASELINE config 4 shape),
  ``interfaces/schur_complement/sc_ip_interface.py:1245-1285``.
* ``dynamic_ipm_system``: time-decomposed layout (BASELINE config 3 shape), ``sc_ip_interface.py:274-357``.
"""
import numpy as np
import scipy.sparse as sp

from parapint_b200.carriers import BlockMatrix, BlockVector  # noqa: F401


def ipm_kkt_block(rng, n_x, n_eq, n_in, n_fs, barrier_span=(1e-4, 1e4), hess_shift=1.0, pattern_rng=None):
    """Parapint-shaped primal-dual KKT of one scenario ("family P"):
    order [x, s, lam_eq, lam_in, lam_link]  (interfaces/interface.py:475-489 wrapped with the linking rows of
    interfaces/schur_complement/sc_ip_interface.py:1248-1266).  Returns (K sparse symmetric, n)."""
    band = sp.diags([rng.standard_normal(n_x - 2) * 0.3, rng.standard_normal(n_x - 1) * 0.5,
                     np.abs(rng.standard_normal(n_x)) + hess_shift,
                     rng.standard_normal(n_x - 1) * 0.0, rng.standard_normal(n_x - 2) * 0.0], [-2, -1, 0, 1, 2]).tocsr()
    H = sp.tril(band) + sp.tril(band, -1).T
    lo, hi = np.log(barrier_span[0]), np.log(barrier_span[1])
    sig_x = np.exp(rng.uniform(lo, hi, n_x))
    sig_s = np.exp(rng.uniform(lo, hi, n_in))
    assert n_fs + n_eq + n_in <= n_x
    # full row rank by construction: the identity parts of L, J_eq and J_in sit on disjoint primal columns
    J_eq = (sp.eye(n_eq, n_x, k=n_fs) + sp.diags([rng.standard_normal(n_eq) * 0.5], [n_fs + 1], shape=(n_eq, n_x))
            + sp.diags([rng.standard_normal(n_eq) * 0.3], [n_fs - 1], shape=(n_eq, n_x))).tocsr()
    if pattern_rng is None:
        R = sp.random(n_in, n_x, density=2.0 / n_x, random_state=rng, data_rvs=rng.standard_normal)
    else:  # every scenario of one model shares the sparsity pattern; only the data differ
        R = sp.random(n_in, n_x, density=2.0 / n_x, random_state=pattern_rng).tocoo()
        R = sp.coo_matrix((rng.standard_normal(R.nnz), (R.row, R.col)), shape=R.shape)
    J_in = (R * 0.3 + sp.eye(n_in, n_x, k=n_x - n_in)).tocsr()
    L = sp.eye(n_fs, n_x).tocsr()  # linking rows select the first n_fs primals
    Z = lambda a, b: sp.csr_matrix((a, b))
    I_in = sp.identity(n_in, format="csr")
    K = sp.bmat([
        [H + sp.diags(sig_x), Z(n_x, n_in), J_eq.T, J_in.T, L.T],
        [Z(n_in, n_x), sp.diags(sig_s), Z(n_in, n_eq), -I_in, Z(n_in, n_fs)],
        [J_eq, Z(n_eq, n_in), Z(n_eq, n_eq), Z(n_eq, n_in), Z(n_eq, n_fs)],
        [J_in, -I_in, Z(n_in, n_eq), Z(n_in, n_in), Z(n_in, n_fs)],
        [L, Z(n_fs, n_in), Z(n_fs, n_eq), Z(n_fs, n_in), Z(n_fs, n_fs)],
    ]).tocoo()
    return K, n_x + n_in + n_eq + n_in + n_fs


def stochastic_ipm_system(seed, n_blocks, n_x, n_eq, n_in, n_fs, same_pattern=False, local_blocks=None, **kw):
    """Block-bordered KKT of a two-stage stochastic NLP: border = [0 | -I] on the linking multipliers
    (sc_ip_interface.py:1275-1280), Q = 0 (:1282-1284)."""
    kkt = BlockMatrix(n_blocks + 1, n_blocks + 1)
    sizes = []
    for i in range(n_blocks):
        n = n_x + 2 * n_in + n_eq + n_fs
        kkt.set_row_size(i, n)
        kkt.set_col_size(i, n)
        if local_blocks is not None and i not in local_blocks:   # another rank's block
            sizes.append(n)
            continue
        rng = np.random.default_rng(1000 * seed + i)
        if same_pattern:
            kw["pattern_rng"] = np.random.default_rng(77 + seed)
        K, n = ipm_kkt_block(rng, n_x, n_eq, n_in, n_fs, **kw)
        sizes.append(n)
        kkt.set_block(i, i, K)
        A = sp.coo_matrix((-np.ones(n_fs), (np.arange(n_fs), n - n_fs + np.arange(n_fs))), shape=(n_fs, n))
        kkt.set_block(n_blocks, i, A)
    kkt.set_block(n_blocks, n_blocks, sp.coo_matrix((n_fs, n_fs)))
    sizes.append(n_fs)
    return kkt, sizes


def dynamic_ipm_system(seed, n_blocks, n_x, n_eq, n_in, n_s, same_pattern=True, local_blocks=None, **kw):
    """Block-bordered KKT of a time-decomposed NLP ("family P, dynamic"; layout of
    interfaces/schur_complement/sc_ip_interface.py:274-357): time block i couples to its neighbours through n_s
    states.  Diagonal block i = [[KKT_i, Lb_i^T],[Lb_i, 0]] (backward-link multipliers inside the block; none for
    block 0); coupling part = [forward-link multipliers of blocks 0..N-2 ; coupling variables z (n_s per interface)];
    border of block i: Lf_i (end states) into the rows of its forward multipliers, -I from its backward multipliers
    into the rows of z_{i-1}; Q = [[0, -I],[-I, 0]].  m_c = 2 n_s (N - 1); S is block tridiagonal.
    Blocks not in ``local_blocks`` are left empty (the other ranks' blocks).  Returns (kkt, sizes)."""
    N = n_blocks
    assert 2 * n_s + n_eq + n_in <= n_x
    nf_tot = n_s * (N - 1)
    m_c = 2 * nf_tot
    kkt = BlockMatrix(N + 1, N + 1)
    sizes = []
    for i in range(N):
        nb = 0 if i == 0 else n_s
        n = n_x + 2 * n_in + n_eq + nb
        sizes.append(n)
        if local_blocks is not None and i not in local_blocks:
            continue
        rng = np.random.default_rng(1000 * seed + i)
        if same_pattern:
            kw["pattern_rng"] = np.random.default_rng(77 + seed)
        K, n_chk = ipm_kkt_block(rng, n_x, n_eq, n_in, nb, **kw)
        assert n_chk == n
        kkt.set_block(i, i, K)
        rows, cols, vals = [], [], []
        if i < N - 1:   # Lf_i: end states = primals [c0, c0 + n_s), disjoint from the identity parts of Lb, J_eq, J_in
            c0 = nb + n_eq
            rows += list(n_s * i + np.arange(n_s))
            cols += list(c0 + np.arange(n_s))
            vals += [1.0] * n_s
        if i > 0:       # -Cb_i^T: backward multipliers (last nb columns) against z_{i-1}
            rows += list(nf_tot + n_s * (i - 1) + np.arange(n_s))
            cols += list(n - nb + np.arange(nb))
            vals += [-1.0] * n_s
        kkt.set_block(N, i, sp.coo_matrix((vals, (rows, cols)), shape=(m_c, n)))
    idx = np.arange(nf_tot)   # both triangles are stored, as the reference's nested Q block holds (1,0) and (0,1)
    Q = sp.coo_matrix((np.concatenate([np.zeros(m_c), -np.ones(nf_tot), -np.ones(nf_tot)]),
                       (np.concatenate([np.arange(m_c), nf_tot + idx, idx]), np.concatenate([np.arange(m_c), idx, nf_tot + idx]))),
                      shape=(m_c, m_c))
    kkt.set_block(N, N, Q)
    for i in range(N):
        kkt.set_row_size(i, sizes[i])
        kkt.set_col_size(i, sizes[i])
    sizes.append(m_c)
    return kkt, sizes
