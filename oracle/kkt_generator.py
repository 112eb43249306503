"""Restatement of the reference's seeded synthetic block-bordered KKT generator.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows reference
``parapint/examples/performance/schur_complement/create_model.py:11-143`` and
``utils.py:24-31`` ("family G" in SURVEY.md 8(d)); the NumPy legacy-RNG call
sequence is reproduced call for call so that matrices are bit-identical to the
reference's (checked against fixtures made by the unmodified generator,
``tests/golden/make_golden.py``).

A parameter-estimation problem: every block ``i`` estimates ``q_i`` from noisy
measurements ``y_i = A q_i``; the first ``n_theta`` entries of every ``q_i`` are
tied to the shared ``theta``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from parapint_b200.carriers import BlockMatrix, BlockVector


def random_banded(n, nnz_per_row):
    """``utils.py:24-31``: unit band of odd width scaled entrywise by N(0, 5^2).

    The reference accumulates ``eye(k=0) + eye(k=1) + eye(k=-1) ...`` which
    yields a canonical CSR matrix (rows ascending, columns ascending within a
    row); the normal draws are applied to ``data`` in exactly that order.
    """
    assert nnz_per_row % 2 == 1
    half = (nnz_per_row - 1) // 2
    rows, cols = [], []
    for i in range(n):
        lo, hi = max(0, i - half), min(n - 1, i + half)
        cols.extend(range(lo, hi + 1))
        rows.extend([i] * (hi - lo + 1))
    data = np.ones(len(rows)) * np.random.normal(loc=0, scale=5, size=len(rows))
    return sp.csr_matrix((data, (np.asarray(rows), np.asarray(cols))), shape=(n, n))


class EstimationBlock:
    """``create_model.py:11-64`` (``Block``)."""

    def __init__(self, n_y, n_q, seed, theta, A, P):
        np.random.seed(seed)
        self.n_theta = theta.size
        self.A = A
        self.P = P
        self.q = np.random.normal(loc=5, scale=2, size=n_q)
        self.q[: self.n_theta] = theta
        self.y_hat = A * self.q
        self.y_hat = self.y_hat + np.random.normal(loc=0, scale=0.01 * np.abs(self.y_hat).max(), size=n_y)

    @property
    def dims(self):
        return self.y_hat.size, self.q.size, self.n_theta

    def kkt(self):
        """``:23-47``: [[2I,0,I,0],[0,0,-A^T,P^T],[I,-A,0,0],[0,P,0,0]] as one COO matrix."""
        n_y, n_q, n_t = self.dims
        eye_y = sp.identity(n_y, format="coo")
        grid = BlockMatrix(4, 4)
        for k, sz in enumerate((n_y, n_q, n_y, n_t)):
            grid.set_row_size(k, sz)
            grid.set_col_size(k, sz)
        grid.set_block(0, 0, 2 * eye_y)
        grid.set_block(0, 2, eye_y)
        grid.set_block(1, 2, -self.A.transpose())
        grid.set_block(1, 3, self.P.transpose())
        grid.set_block(2, 0, eye_y)
        grid.set_block(2, 1, -self.A)
        grid.set_block(3, 1, self.P)
        return grid.tocoo()

    def rhs(self):
        """``:49-58``."""
        n_y, n_q, n_t = self.dims
        return np.concatenate([2 * self.y_hat, np.zeros(n_q + n_y + n_t)])

    def error(self, sol):
        """``:60-64``."""
        n_y, n_q, _ = self.dims
        return np.abs(np.asarray(sol)[n_y:n_y + n_q] - self.q).max()


class EstimationModel:
    """``create_model.py:67-143`` (``Model``); ``local_blocks`` mirrors ``MPIModel`` ``:146-255``."""

    def __init__(self, n_blocks, n_q_per_block, n_y_multiplier, n_theta, A_nnz_per_row=3, local_blocks=None):
        assert isinstance(n_y_multiplier, int) and n_y_multiplier > 1
        self.n_blocks = n_blocks
        self.n_q = n_q_per_block
        self.n_y = n_q_per_block * n_y_multiplier
        self.n_theta = n_theta
        np.random.seed(0)
        np.random.seed(np.random.randint(low=0, high=1000000))
        stack = [random_banded(n_q_per_block, A_nnz_per_row) for _ in range(n_y_multiplier)]
        self.A = sp.vstack(stack).tocoo()
        self.theta = np.random.normal(loc=5, scale=2, size=n_theta)
        self.P = sp.coo_matrix((np.ones(n_theta), (np.arange(n_theta), np.arange(n_theta))),
                               shape=(n_theta, n_q_per_block))
        self.local_blocks = list(range(n_blocks)) if local_blocks is None else list(local_blocks)
        self.blocks = {i: EstimationBlock(self.n_y, self.n_q, i, self.theta, self.A, self.P)
                       for i in self.local_blocks}

    @property
    def block_dim(self):
        return 2 * self.n_y + self.n_q + self.n_theta

    def border(self):
        """``:105-110``: ``[0 | -I_theta]`` on the last ``n_theta`` columns of a block."""
        n, t = self.block_dim, self.n_theta
        return sp.coo_matrix((-np.ones(t), (np.arange(t), n - t + np.arange(t))), shape=(t, n))

    def build_kkt(self):
        """``:104-123``: block-bordered KKT, lower and upper border both set, Q empty."""
        N = self.n_blocks
        border = self.border()
        kkt = BlockMatrix(N + 1, N + 1)
        for i in range(N):
            kkt.set_row_size(i, self.block_dim)
            kkt.set_col_size(i, self.block_dim)
        for i in self.local_blocks:
            kkt.set_block(i, i, self.blocks[i].kkt())
            kkt.set_block(N, i, border)
            kkt.set_block(i, N, border.transpose().tocoo())
        kkt.set_block(N, N, sp.coo_matrix((self.n_theta, self.n_theta)))
        return kkt

    def build_rhs(self):
        """``:125-132``."""
        N = self.n_blocks
        rhs = BlockVector(N + 1)
        for i in self.local_blocks:
            rhs.set_block(i, self.blocks[i].rhs())
        rhs.set_block(N, np.zeros(self.n_theta))
        return rhs

    def check_result(self, sol):
        """``:134-143``: max abs error of the estimated q and theta."""
        err = 0.0
        for i in self.local_blocks:
            err = max(err, self.blocks[i].error(sol.get_block(i)))
        return max(err, np.abs(np.asarray(sol.get_block(self.n_blocks)) - self.theta).max())

    def expected_inertia(self):
        """Closed form verified against eigvalsh / dsytrf / the reference."""
        N = self.n_blocks
        return N * (self.n_y + self.n_q) + self.n_theta, N * (self.n_y + self.n_theta), 0
