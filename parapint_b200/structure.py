"""Host-side symbolic analysis of a block-bordered KKT matrix.

Turns the PyNumero-style ``BlockMatrix`` the IPM hands to the linear solver
(layout: reference ``explicit_schur_complement.py:17-27``; inputs described in
SURVEY.md 3.6 / 8(b)) into the flat description the C ABI takes
(``include/parapint_b200.h``, ``pp_symbolic``): which diagonal blocks this rank
owns, the nonzero border rows of every owned block, and for every numeric input
value the front position it is added into.

Ownership follows ``mpi_explicit_schur_complement.py:198-203``: block ``i`` is
local when ``rank_ownership[i, i] == rank`` or (``== -1`` and ``rank == 0``); a
matrix without ``rank_ownership`` (serial ``BlockMatrix``) is split round-robin
``i % size == rank`` (``mpi_sc_ip_interface.py:14-19``) when ``size > 1``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import operator

import numpy as np
import scipy.sparse as sp

try:   # identity comparison of two lists in one call (csrc/fastptr.c); the interpreter does the same, slower
    from . import _pp_fastptr as _fp
except ImportError:
    _fp = None


def _coo(block):
    """COO triplets of a leaf / nested block without summing duplicates (index arrays are not copied)."""
    c = block.tocoo()
    return c.row, c.col, np.asarray(c.data, dtype=np.float64), c.shape


def _is_nested(block) -> bool:
    return hasattr(block, "bshape") and hasattr(block, "get_block")


_KIND_BY_TYPE = {}   # class -> 0 (COO leaf) / 1 (nested block matrix) / 2 (anything else)


def _kind(sub):
    """0 for a COO leaf, 1 for a nested block matrix, 2 otherwise.  For SciPy's sparse classes and for block-matrix
    classes the answer is a property of the class and is remembered per class (``format`` is a Python-level property
    in SciPy: asking ~30 sub-leaves per scenario and iteration costs as much as walking them)."""
    cls = type(sub)
    kind = _KIND_BY_TYPE.get(cls)
    if kind is None:
        if _is_nested(sub):
            kind = _KIND_BY_TYPE[cls] = 1
        elif sp.issparse(sub):
            kind = _KIND_BY_TYPE[cls] = 0 if sub.format == "coo" else 2
        else:   # a duck-typed leaf: its format may be an instance attribute, ask every time
            kind = 0 if getattr(sub, "format", None) == "coo" else 2
    return kind


def _walk_leaves(block, leaves, sig):
    """COO sub-leaves of a nested block matrix in row-major block order (the order ``tocoo()`` concatenates them),
    appended to ``leaves``; ``sig`` receives the shape of the walk as a flat list of integers (position of every
    leaf, entry / exit marks of every nested sub-block), so that two walks visited the same positions exactly when
    their signatures are equal.  Returns False when a sub-block is neither nested nor a COO matrix."""
    nbr, nbc = block.bshape
    get = block.get_block
    kinds = _KIND_BY_TYPE
    pos = 0
    for i in range(nbr):
        for j in range(nbc):
            sub = get(i, j)
            if sub is not None:
                kind = kinds.get(type(sub))
                if kind is None:
                    kind = _kind(sub)
                if kind == 0:
                    leaves.append(sub)
                    sig.append(pos)
                elif kind == 1:
                    sig.append(-1 - pos)
                    if not _walk_leaves(sub, leaves, sig):
                        return False
                    sig.append(-1 - nbr * nbc)
                else:
                    return False
            pos += 1
    return True


def _leaf_index(leaf):
    """(row, col) index arrays of a COO leaf without copying (scipy's ``row`` / ``col`` are properties of ``coords``)."""
    coords = getattr(leaf, "coords", None)
    return coords if coords is not None else (leaf.row, leaf.col)


def _flatten_recipe(block, data):
    """How to gather the values of a nested leaf without flattening it every iteration (parapint builds a new
    nested 4x4 BlockMatrix per scenario and iteration, ``interface.py:432-491``; its ``tocoo()`` costs far more
    than the factorisation): the walk's signature and the index arrays of the COO sub-leaves in concatenation order.
    The recipe is only kept when concatenating the sub-leaves' values reproduces ``tocoo().data`` exactly."""
    if not _is_nested(block):
        return None
    leaves, sig = [], []
    if not _walk_leaves(block, leaves, sig) or not leaves:
        return None
    parts = [np.asarray(leaf.data, dtype=np.float64) for leaf in leaves]
    if sum(p.size for p in parts) != data.size or not np.array_equal(np.concatenate(parts), data):
        return None
    index = [_leaf_index(leaf) for leaf in leaves]
    sizes = [p.size for p in parts]
    rel = [0]
    for n in sizes[:-1]:
        rel.append(rel[-1] + n)
    return sig, [ix[0] for ix in index], [ix[1] for ix in index], sizes, rel


@dataclass
class Structure:
    n_blocks: int                      # N (number of diagonal blocks, all ranks)
    m_c: int                           # coupling dimension
    local_blocks: List[int]            # global indices of the blocks owned here
    block_n: np.ndarray                # int32[n_local]
    border_ptr: np.ndarray             # int64[n_local + 1]
    border_rows: np.ndarray            # int32[sum m_i]
    dest_front: np.ndarray             # int32[nvals]
    dest_row: np.ndarray               # int32[nvals]
    dest_col: np.ndarray               # int32[nvals]
    segments: list = field(default_factory=list)   # (kind, block index, start, stop) per gathered COO array
    patterns: list = field(default_factory=list)   # (row, col) arrays per segment, to detect pattern changes
    recipes: list = field(default_factory=list)    # per segment: sub-leaf recipe of a nested block, or None
    rhs_offsets: np.ndarray = None     # int64[n_local + 1] offsets of local blocks in the packed rhs

    @property
    def nvals(self):
        return int(self.dest_front.size)

    @property
    def n_local(self):
        return len(self.local_blocks)

    @property
    def segment_starts(self):
        starts = self.__dict__.get("_segment_starts")
        if starts is None:
            starts = self.__dict__["_segment_starts"] = [seg[2] for seg in self.segments]
        return starts

    @property
    def local_dim(self):
        return int(self.rhs_offsets[-1])


def local_block_indices(matrix, n_blocks, rank, size):
    own = getattr(matrix, "rank_ownership", None)
    if own is not None:
        own = np.asarray(own)
        return [i for i in range(n_blocks) if own[i, i] == rank or (own[i, i] == -1 and rank == 0)]
    return [i for i in range(n_blocks) if i % size == rank]


def analyse(matrix, rank=0, size=1) -> Structure:
    nbr, nbc = matrix.bshape
    if nbr != nbc:
        raise ValueError("The block matrix provided is not square.")
    if getattr(matrix, "rank_ownership", None) is None:
        nrows, ncols = matrix.shape
        if nrows != ncols:
            raise ValueError("The block matrix provided is not square.")
    N = nbr - 1
    local = local_block_indices(matrix, N, rank, size)
    Q = matrix.get_block(N, N)
    if Q is None:
        m_c = int(matrix.get_row_size(N))
        q_row = q_col = np.zeros(0, dtype=np.int64)
    else:
        q_row, q_col, _, qshape = _coo(Q)
        m_c = int(qshape[0])
        if qshape[0] != qshape[1]:
            raise ValueError("The coupling block is not square.")

    block_n, border_ptr, border_rows = [], [0], []
    dfront, drow, dcol = [], [], []
    segments, patterns, recipes = [], [], []
    pos = 0
    for f, i in enumerate(local):
        K = matrix.get_block(i, i)
        if K is None:
            raise ValueError(f"diagonal block {i} is missing on its owner")
        kr, kc, kdata, kshape = _coo(K)
        if kshape[0] != kshape[1]:
            raise ValueError(f"diagonal block {i} is not square")
        n_i = int(kshape[0])
        block_n.append(n_i)
        keep = kr >= kc  # lower triangle only, like the MA27 / MUMPS leaves (mumps_interface.py:51)
        dfront.append(np.where(keep, f, -1))
        drow.append(np.where(keep, kr, 0))
        dcol.append(np.where(keep, kc, 0))
        segments.append(("K", i, pos, pos + kr.size))
        patterns.append((kr, kc))
        recipes.append(_flatten_recipe(K, kdata))
        pos += kr.size

        A = matrix.get_block(N, i)
        if A is None:
            border_ptr.append(border_ptr[-1])
            continue
        ar, ac, adata, ashape = _coo(A)
        if ashape != (m_c, n_i):
            raise ValueError(f"border block ({N},{i}) has shape {ashape}, expected {(m_c, n_i)}")
        nz_rows = np.unique(ar)  # rows with stored entries (explicit zeros count, as in _BorderMatrix)
        lookup = np.full(m_c, -1, dtype=np.int64)
        lookup[nz_rows] = np.arange(nz_rows.size)
        border_rows.append(nz_rows)
        border_ptr.append(border_ptr[-1] + nz_rows.size)
        dfront.append(np.full(ar.size, f, dtype=np.int64))
        drow.append(n_i + lookup[ar])
        dcol.append(ac)
        segments.append(("A", i, pos, pos + ar.size))
        patterns.append((ar, ac))
        recipes.append(_flatten_recipe(A, adata))
        pos += ar.size

    n_local = len(local)
    if Q is not None:
        keep = q_row >= q_col
        dfront.append(np.where(keep, n_local, -1))
        drow.append(np.where(keep, q_row, 0))
        dcol.append(np.where(keep, q_col, 0))
        segments.append(("Q", N, pos, pos + q_row.size))
        patterns.append((q_row, q_col))
        recipes.append(None)
        pos += q_row.size

    def cat(parts, dtype):
        return np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0), dtype=dtype)

    offs = np.concatenate(([0], np.cumsum(block_n))).astype(np.int64) if block_n else np.zeros(1, dtype=np.int64)
    return Structure(
        n_blocks=N, m_c=m_c, local_blocks=local,
        block_n=np.asarray(block_n, dtype=np.int32),
        border_ptr=np.asarray(border_ptr, dtype=np.int64),
        border_rows=cat(border_rows, np.int32),
        dest_front=cat(dfront, np.int32), dest_row=cat(drow, np.int32), dest_col=cat(dcol, np.int32),
        segments=segments, patterns=patterns, recipes=recipes, rhs_offsets=offs)


def _same_objects(a, b) -> bool:
    """Both lists hold the same objects (identity, not equality)."""
    if _fp is not None:
        return _fp.same(a, b)
    return len(a) == len(b) and all(map(operator.is_, a, b))


def _same_index(a, b) -> bool:
    if a is b:
        return True
    if a.size != b.size:
        return False
    if a.dtype == b.dtype and a.flags.c_contiguous and b.flags.c_contiguous:
        return a.tobytes() == b.tobytes()   # 4x cheaper than np.array_equal on leaves of a few thousand entries
    return bool(np.array_equal(a, b))


def gather_values(matrix, st: Structure, out: np.ndarray, copier=None) -> bool:
    """Copy the numeric values of ``matrix`` into ``out`` in the order fixed by :func:`analyse`.

    Returns False when a block's COO pattern differs from the analysed one (the caller then
    re-runs the symbolic phase, as ``mumps_interface.py:82-83`` does).  ``copier`` (a
    ``native.HostCopier``) moves the leaves with a few threads instead of one numpy slice assignment each.

    Three paths per leaf, fastest first: the very object validated last time (values updated in place); a nested
    block matrix walked sub-leaf by sub-leaf against the recipe recorded by :func:`analyse` (index arrays compared
    unless the sub-leaves are the objects validated last time, nothing concatenated); ``tocoo()`` and a comparison
    of the flattened index arrays.  All index comparisons of a call are made together at the end, by the copier's
    threads when there is one."""
    N = st.n_blocks
    get = matrix.get_block
    # whole-matrix fast path: every segment is the flat COO leaf validated last time and carries the index tuple it had
    # then (an interior-point loop that updates the values in place): four list passes, no per-leaf branching
    fast = st.__dict__.get("_fast_leaves")
    if fast is not None:
        blks = [get(i, j) for i, j in fast[0]]
        if _same_objects(blks, fast[1]) and _same_objects([b.coords for b in blks], fast[2]):
            datas = [b.data for b in blks]
            if [d.size for d in datas] != fast[3]:
                return False
            starts = st.segment_starts
            if copier is None or not copier.copy(datas, starts, out):
                for lo, data in zip(starts, datas):
                    out[lo:lo + data.size] = data
            return True
    datas, starts = [], []
    fresh_idx, ref_idx = [], []   # fresh / analysed index arrays, compared in one threaded call at the end
    validated = []                # (segment, leaf object(s), index tuple(s)) that pass once that comparison has passed
    seen = st.__dict__.setdefault("_leaf_seen", [None] * len(st.segments))
    recipes = st.recipes if st.recipes else [None] * len(st.segments)
    for k, ((kind, i, lo, hi), (prow, pcol)) in enumerate(zip(st.segments, st.patterns)):
        blk = get(i, i) if kind != "A" else get(N, i)
        if blk is None:
            return False
        # fastest path: the very leaf object validated last time, index tuple untouched (values updated in place)
        last = seen[k]
        if last is not None and last[0] is blk and getattr(blk, "coords", None) is last[1]:
            data = blk.data
            if data.size != hi - lo:
                return False
            datas.append(data)
            starts.append(lo)
            continue
        recipe = recipes[k]
        kind_blk = _KIND_BY_TYPE.get(type(blk))
        if kind_blk is None:
            kind_blk = _kind(blk)
        if recipe is not None and kind_blk == 1:
            leaves, sig = [], []
            if not _walk_leaves(blk, leaves, sig) or sig != recipe[0]:
                return False
            coords = [getattr(leaf, "coords", None) for leaf in leaves]
            if None in coords:
                coords = [(leaf.row, leaf.col) for leaf in leaves]
            # the sub-leaves validated last time, index tuples untouched: nothing to compare
            if last is None or not (_same_objects(leaves, last[0]) and _same_objects(coords, last[1])):
                for (lrow, lcol), rrow, rcol in zip(coords, recipe[1], recipe[2]):
                    if lrow is not rrow:
                        fresh_idx.append(lrow)
                        ref_idx.append(rrow)
                    if lcol is not rcol:
                        fresh_idx.append(lcol)
                        ref_idx.append(rcol)
                seen[k] = None
                validated.append((k, leaves, coords))
            part = [leaf.data for leaf in leaves]
            if [d.size for d in part] != recipe[3]:
                return False
            datas.extend(part)
            starts.extend([lo + r for r in recipe[4]] if lo else recipe[4])
            continue
        # (scipy's row / col / format are properties: every access costs; the index tuple is read once per leaf)
        is_coo = kind_blk == 0
        c = blk if is_coo else blk.tocoo()
        coords = getattr(c, "coords", None)
        frow, fcol = coords if coords is not None else (c.row, c.col)
        same = frow is prow and fcol is pcol      # still carries the analysed index arrays (values updated in place)
        if not same:
            # (sizes and types are checked by the comparison at the end)
            if frow is not prow:
                fresh_idx.append(frow)
                ref_idx.append(prow)
            if fcol is not pcol:
                fresh_idx.append(fcol)
                ref_idx.append(pcol)
        data = c.data
        if data.size != hi - lo:
            return False
        seen[k] = None
        if is_coo and coords is not None:
            validated.append((k, blk, coords))
        datas.append(data)
        starts.append(lo)
    if fresh_idx:
        same = copier.all_equal(fresh_idx, ref_idx) if copier is not None else None
        if same is None:   # no threaded comparison, or a pair of different types / lengths
            same = all(map(_same_index, fresh_idx, ref_idx))
        if not same:
            return False
    for k, obj, coords in validated:   # these objects now count as validated while their index tuples stay untouched
        seen[k] = (obj, coords)
    flat = [sk for sk in seen if sk is not None and type(sk[1]) is tuple]
    if len(flat) == len(seen):   # every segment is a flat leaf with an index tuple: remember them for the fast path
        keys = [(i, i) if kind != "A" else (N, i) for kind, i, _, _ in st.segments]
        st.__dict__["_fast_leaves"] = (keys, [sk[0] for sk in flat], [sk[1] for sk in flat],
                                       [hi - lo for _, _, lo, hi in st.segments])
    else:
        st.__dict__.pop("_fast_leaves", None)
    if copier is None or not copier.copy(datas, starts, out):
        for lo, data in zip(starts, datas):
            out[lo:lo + data.size] = data
    return True


def pack_rhs(rhs, st: Structure, out: np.ndarray, copier=None):
    if copier is not None:
        get = rhs.get_block
        blocks = [get(i) for i in st.local_blocks]
        sizes = st.__dict__.get("_rhs_sizes")
        if sizes is None:
            sizes = st.__dict__["_rhs_sizes"] = [int(n) for n in np.diff(st.rhs_offsets)]
        if {type(b) for b in blocks} <= {np.ndarray} and [b.size for b in blocks] == sizes \
                and copier.copy(blocks, st.rhs_offsets[:-1], out):   # (the copier insists on contiguous float64)
            return
    for f, i in enumerate(st.local_blocks):
        blk = rhs.get_block(i)
        flat = blk.flatten() if hasattr(blk, "nblocks") else np.asarray(blk, dtype=np.float64).ravel()
        lo, hi = st.rhs_offsets[f], st.rhs_offsets[f + 1]
        if flat.size != hi - lo:
            raise ValueError(f"rhs block {i} has size {flat.size}, expected {hi - lo}")
        out[lo:hi] = flat


def coupling_rhs(rhs, st: Structure) -> np.ndarray:
    blk = rhs.get_block(st.n_blocks)
    flat = blk.flatten() if hasattr(blk, "nblocks") else np.asarray(blk, dtype=np.float64).ravel()
    if flat.size != st.m_c:
        raise ValueError(f"coupling rhs has size {flat.size}, expected {st.m_c}")
    return np.ascontiguousarray(flat, dtype=np.float64)


def _views_like(template, seg):
    """A vector with the (nested) block structure of ``template`` whose leaves are views of ``seg`` -- one pass, no
    zero-filled intermediate as ``copy_structure()`` + ``copyfrom()`` would need."""
    if not hasattr(template, "nblocks"):
        return seg
    out = type(template)(template.nblocks)
    pos = 0
    for k in range(template.nblocks):
        sub = template.get_block(k)
        n = int(sub.size)
        out.set_block(k, _views_like(sub, seg[pos:pos + n]))
        pos += n
    if pos != seg.size:
        raise ValueError("block sizes do not add up")
    return out


def _nested_like(template, values):
    try:
        return _views_like(template, np.array(values, dtype=np.float64))
    except Exception:  # noqa: BLE001 - an unusual vector class: go through its own copy_structure / copyfrom
        blk = template.copy_structure()
        blk.copyfrom(values)
        return blk


THREADED_UNPACK_BYTES = 4 << 20   # solutions of at least this size leave the pinned buffer through the copy pool ...
AWAKE_UNPACK_BYTES = 256 << 10    # ... or of this size when the pool has been told to stay awake for them (copier.wake_us)


def unpack_solution(rhs, st: Structure, x_local: np.ndarray, x_c: np.ndarray, copier=None):
    """New vector with the block structure of ``rhs`` (``mpi_explicit_schur_complement.py:390``;
    nested blocks keep their structure as the SciPy leaf does, ``scipy_interface.py:57-60``).
    The blocks are views of one freshly allocated array (nothing aliases solver buffers); ``copier`` (a
    ``native.HostCopier``) fills that array with a few threads when the solution is large (20 MB per rank at
    BASELINE configs 3 and 4: 4 ms for one thread)."""
    out = rhs.empty_like_structure() if hasattr(rhs, "empty_like_structure") else rhs.copy_structure()
    n = st.local_dim
    flat = None
    limit = THREADED_UNPACK_BYTES
    if copier is not None and getattr(copier, "wake_us", 0) > 0 and getattr(copier, "threads", 1) > 1:
        limit = min(limit, AWAKE_UNPACK_BYTES)
    if copier is not None and n * 8 >= limit and x_local.dtype == np.float64 \
            and x_local.flags.c_contiguous:
        flat = np.empty(n, dtype=np.float64)
        if not copier.copy([flat], [0], x_local, to_staging=False):
            flat = None
    if flat is None:
        flat = np.array(x_local[:n], dtype=np.float64)  # one copy out of the pinned buffer
    bounds = st.__dict__.get("_rhs_bounds")
    if bounds is None:   # plain integers: slicing with NumPy scalars costs a conversion per bound
        offs = [int(v) for v in st.rhs_offsets]
        bounds = st.__dict__["_rhs_bounds"] = list(zip(offs[:-1], offs[1:]))
    get, put = rhs.get_block, out.set_block
    ndarray = np.ndarray
    for i, (lo, hi) in zip(st.local_blocks, bounds):
        template = get(i)
        seg = flat[lo:hi]
        if type(template) is ndarray or not hasattr(template, "nblocks"):
            put(i, seg)
        else:
            try:
                put(i, _views_like(template, seg))
            except Exception:  # noqa: BLE001 - an unusual vector class
                blk = template.copy_structure()
                blk.copyfrom(seg)
                put(i, blk)
    template = get(st.n_blocks)
    if hasattr(template, "nblocks"):
        put(st.n_blocks, _nested_like(template, x_c))
    else:
        put(st.n_blocks, np.array(x_c, dtype=np.float64))
    return out
