"""Shared builders for the parity tests (inputs only; no solver logic)."""
import numpy as np
import scipy.sparse as sp

from parapint_b200.carriers import BlockMatrix, BlockVector


def bordered_from_dense(dense, block_sizes, lower_only=True):
    """Cut a dense symmetric block-bordered matrix into a BlockMatrix (last size = coupling)."""
    nb = len(block_sizes)
    off = np.concatenate(([0], np.cumsum(block_sizes)))
    kkt = BlockMatrix(nb, nb)
    for i in range(nb):
        kkt.set_block(i, i, sp.coo_matrix(dense[off[i]:off[i + 1], off[i]:off[i + 1]]))
    for i in range(nb - 1):
        kkt.set_block(nb - 1, i, sp.coo_matrix(dense[off[nb - 1]:, off[i]:off[i + 1]]))
        if not lower_only:
            kkt.set_block(i, nb - 1, sp.coo_matrix(dense[off[i]:off[i + 1], off[nb - 1]:]))
    return kkt


def block_vector(flat, block_sizes):
    off = np.concatenate(([0], np.cumsum(block_sizes)))
    v = BlockVector(len(block_sizes))
    for i in range(len(block_sizes)):
        v.set_block(i, np.array(flat[off[i]:off[i + 1]], dtype=np.float64))
    return v


def random_bordered(rng, n_blocks, n, m_c, density=0.05, border_nnz_rows=None, definite_shift=0.0):
    """Random sparse symmetric-indefinite block-bordered system with a general (non-selection) border."""
    kkt = BlockMatrix(n_blocks + 1, n_blocks + 1)
    for i in range(n_blocks):
        ni = n if np.isscalar(n) else n[i]
        M = sp.random(ni, ni, density=density, random_state=rng, data_rvs=rng.standard_normal).toarray()
        K = M + M.T + np.diag(rng.standard_normal(ni) * 2.0) + definite_shift * np.eye(ni)
        kkt.set_block(i, i, sp.coo_matrix(K))
        rows = m_c if border_nnz_rows is None else border_nnz_rows
        A = np.zeros((m_c, ni))
        pick = rng.choice(m_c, size=min(rows, m_c), replace=False)
        A[pick] = sp.random(len(pick), ni, density=min(1.0, 4.0 / ni + density), random_state=rng,
                            data_rvs=rng.standard_normal).toarray()
        kkt.set_block(n_blocks, i, sp.coo_matrix(A))
    Qh = rng.standard_normal((m_c, m_c))
    kkt.set_block(n_blocks, n_blocks, sp.coo_matrix(Qh + Qh.T))
    return kkt


from oracle.kkt_families import dynamic_ipm_system, ipm_kkt_block, stochastic_ipm_system  # noqa: E402,F401
