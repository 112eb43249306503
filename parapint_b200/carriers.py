"""Block-structured matrix / vector carriers.

parapint hands its linear solvers PyNumero ``BlockMatrix`` / ``BlockVector``
objects (external package ``pyomo.contrib.pynumero.sparse``; call sites listed
in SURVEY.md Appendix A, e.g. reference
``parapint/linalg/schur_complement/explicit_schur_complement.py:60,73,108,115``
and ``parapint/linalg/scipy_interface.py:50-60``).  The solver in this package
only duck-types on that surface (``bshape``, ``get_block``, ``tocoo`` ...), so a
real PyNumero object works unchanged.  When Pyomo is not installed (it is not in
this image) these light-weight stand-ins provide the same surface so that the
solver, the tests and the benchmark can build block-bordered systems.

Only host-side bookkeeping lives here; no arithmetic of the hot path does.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = ["BlockMatrix", "BlockVector", "is_block_matrix", "is_block_vector"]


def is_block_matrix(obj) -> bool:
    return hasattr(obj, "bshape") and hasattr(obj, "get_block") and not hasattr(obj, "nblocks")


def is_block_vector(obj) -> bool:
    return hasattr(obj, "nblocks") and hasattr(obj, "get_block")


def _as_coo(block):
    """COO view of a leaf or nested block (duplicates are preserved)."""
    if is_block_matrix(block):
        return block.tocoo()
    if sp.issparse(block):
        return block.tocoo()
    return sp.coo_matrix(np.asarray(block, dtype=np.float64))


class BlockMatrix:
    """Grid of optional sparse blocks with explicit row / column block sizes."""

    def __init__(self, nbrows: int, nbcols: int):
        self._nbrows = int(nbrows)
        self._nbcols = int(nbcols)
        self._blocks = [[None] * self._nbcols for _ in range(self._nbrows)]
        self._row_sizes = [None] * self._nbrows
        self._col_sizes = [None] * self._nbcols

    # ---- structure ---------------------------------------------------------
    @property
    def bshape(self):
        return self._nbrows, self._nbcols

    @property
    def shape(self):
        return sum(self._need_sizes(self._row_sizes, "row")), sum(self._need_sizes(self._col_sizes, "col"))

    @staticmethod
    def _need_sizes(sizes, what):
        if any(s is None for s in sizes):
            raise RuntimeError(f"BlockMatrix has undefined {what} block sizes")
        return sizes

    def get_row_size(self, i):
        return self._row_sizes[i]

    def get_col_size(self, j):
        return self._col_sizes[j]

    def set_row_size(self, i, n):
        self._row_sizes[i] = int(n)

    def set_col_size(self, j, n):
        self._col_sizes[j] = int(n)

    def row_block_sizes(self, copy=True):
        return np.asarray(self._need_sizes(self._row_sizes, "row"), dtype=np.int64)

    def col_block_sizes(self, copy=True):
        return np.asarray(self._need_sizes(self._col_sizes, "col"), dtype=np.int64)

    def is_empty_block(self, i, j):
        return self._blocks[i][j] is None

    def get_block(self, i, j):
        return self._blocks[i][j]

    def set_block(self, i, j, block):
        if block is None:
            self._blocks[i][j] = None
            return
        if not (is_block_matrix(block) or sp.issparse(block)):
            block = sp.coo_matrix(np.asarray(block, dtype=np.float64))
        nr, nc = block.shape
        for sizes, k, n, what in ((self._row_sizes, i, nr, "row"), (self._col_sizes, j, nc, "col")):
            if sizes[k] is None:
                sizes[k] = int(n)
            elif sizes[k] != n:
                raise ValueError(f"block ({i},{j}) has {n} {what}s, expected {sizes[k]}")
        self._blocks[i][j] = block

    # ---- conversion --------------------------------------------------------
    def tocoo(self):
        nrows, ncols = self.shape
        roff = np.concatenate(([0], np.cumsum(self._row_sizes)))
        coff = np.concatenate(([0], np.cumsum(self._col_sizes)))
        rows, cols, vals = [], [], []
        for i in range(self._nbrows):
            for j in range(self._nbcols):
                blk = self._blocks[i][j]
                if blk is None:
                    continue
                c = _as_coo(blk)
                rows.append(c.row.astype(np.int64) + roff[i])
                cols.append(c.col.astype(np.int64) + coff[j])
                vals.append(np.asarray(c.data, dtype=np.float64))
        if rows:
            rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        else:
            rows = cols = np.zeros(0, dtype=np.int64)
            vals = np.zeros(0)
        return sp.coo_matrix((vals, (rows, cols)), shape=(nrows, ncols))

    def tocsr(self):
        return self.tocoo().tocsr()

    def tocsc(self):
        return self.tocoo().tocsc()

    def toarray(self):
        return self.tocoo().toarray()

    def transpose(self, axes=None, copy=True):
        out = BlockMatrix(self._nbcols, self._nbrows)
        out._row_sizes = list(self._col_sizes)
        out._col_sizes = list(self._row_sizes)
        for i in range(self._nbrows):
            for j in range(self._nbcols):
                blk = self._blocks[i][j]
                if blk is not None:
                    out._blocks[j][i] = blk.transpose(copy=True) if is_block_matrix(blk) else blk.transpose().tocoo()
        return out

    def copy(self, deep=True):
        out = self.copy_structure()
        for i in range(self._nbrows):
            for j in range(self._nbcols):
                blk = self._blocks[i][j]
                if blk is not None:
                    out._blocks[i][j] = blk.copy()
        return out

    def copy_structure(self):
        out = BlockMatrix(self._nbrows, self._nbcols)
        out._row_sizes = list(self._row_sizes)
        out._col_sizes = list(self._col_sizes)
        return out

    # ---- minimal arithmetic (host convenience; used by tests only) ----------
    def __add__(self, other):
        if not is_block_matrix(other) or other.bshape != self.bshape:
            return NotImplemented
        out = self.copy_structure()
        for i in range(self._nbrows):
            for j in range(self._nbcols):
                a, b = self._blocks[i][j], other.get_block(i, j)
                if a is None and b is None:
                    continue
                if a is None:
                    out.set_block(i, j, b.copy())
                elif b is None:
                    out.set_block(i, j, a.copy())
                else:
                    out.set_block(i, j, (_as_coo(a) + _as_coo(b)).tocoo())
        return out

    def __mul__(self, other):
        if np.isscalar(other):
            out = self.copy()
            for i in range(self._nbrows):
                for j in range(self._nbcols):
                    if out._blocks[i][j] is not None:
                        out._blocks[i][j] = out._blocks[i][j] * other
            return out
        vec = other.flatten() if is_block_vector(other) else np.asarray(other)
        return self.tocsr().dot(vec)

    dot = __mul__


class BlockVector:
    """Sequence of dense blocks (ndarray or nested ``BlockVector``)."""

    __array_priority__ = 100.0

    def __init__(self, nblocks: int):
        self._blocks = [None] * int(nblocks)

    @property
    def nblocks(self):
        return len(self._blocks)

    @property
    def bshape(self):
        return (len(self._blocks),)

    @property
    def size(self):
        return int(sum(b.size for b in self._blocks))

    @property
    def shape(self):
        return (self.size,)

    @property
    def ndim(self):
        return 1

    @property
    def dtype(self):
        return np.dtype(np.float64)

    def block_sizes(self, copy=True):
        return np.asarray([b.size for b in self._blocks], dtype=np.int64)

    def get_block(self, i):
        return self._blocks[i]

    def set_block(self, i, block):
        if not is_block_vector(block):
            block = np.asarray(block, dtype=np.float64)
        self._blocks[i] = block

    def flatten(self, order="C"):
        if not self._blocks:
            return np.zeros(0)
        parts = [b.flatten() if is_block_vector(b) else np.asarray(b, dtype=np.float64).ravel() for b in self._blocks]
        return np.concatenate(parts)

    def __array__(self, dtype=None, copy=None):
        out = self.flatten()
        return out if dtype is None else out.astype(dtype)

    def copy(self, order="C"):
        out = BlockVector(self.nblocks)
        for i, b in enumerate(self._blocks):
            out._blocks[i] = None if b is None else b.copy()
        return out

    def copy_structure(self):
        out = BlockVector(self.nblocks)
        for i, b in enumerate(self._blocks):
            if b is None:
                continue
            out._blocks[i] = b.copy_structure() if is_block_vector(b) else np.zeros(b.size, dtype=np.float64)
        return out

    def empty_like_structure(self):
        """Same number of blocks, nothing allocated (the caller sets every block it owns)."""
        return BlockVector(self.nblocks)

    def copyfrom(self, other):
        flat = other.flatten() if is_block_vector(other) else np.asarray(other, dtype=np.float64).ravel()
        if flat.size != self.size:
            raise ValueError("size mismatch in BlockVector.copyfrom")
        pos = 0
        for i, b in enumerate(self._blocks):
            n = b.size
            if is_block_vector(b):
                b.copyfrom(flat[pos:pos + n])
            else:
                self._blocks[i] = flat[pos:pos + n].copy()
            pos += n

    def fill(self, value):
        for b in self._blocks:
            b.fill(value)

    # ---- arithmetic preserving the block structure ---------------------------
    def _binary(self, other, op):
        out = self.copy_structure()
        if np.isscalar(other):
            out.copyfrom(op(self.flatten(), other))
        else:
            rhs = other.flatten() if is_block_vector(other) else np.asarray(other, dtype=np.float64).ravel()
            out.copyfrom(op(self.flatten(), rhs))
        return out

    def __add__(self, other):
        return self._binary(other, np.add)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary(other, np.subtract)

    def __rsub__(self, other):
        return self._binary(other, lambda a, b: b - a)

    def __mul__(self, other):
        return self._binary(other, np.multiply)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self._binary(other, np.divide)

    def __neg__(self):
        return self._binary(-1.0, np.multiply)

    def __isub__(self, other):
        self.copyfrom(self._binary(other, np.subtract))
        return self

    def __iadd__(self, other):
        self.copyfrom(self._binary(other, np.add))
        return self

    def __len__(self):
        return self.size

    def max(self):
        return self.flatten().max()

    def min(self):
        return self.flatten().min()

    def __abs__(self):
        return self._binary(1.0, lambda a, b: np.abs(a))

    def __repr__(self):
        return f"BlockVector({self.nblocks} blocks, size {self.size})"
