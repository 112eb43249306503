"""ctypes binding of the C ABI declared in ``include/parapint_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU implementation behind this module: if the library is missing or no CUDA device is present the
solver raises, it never falls back.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libparapint_b200.so")

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

#: every symbol the header declares: name -> (restype, argtypes)
SIGNATURES = {
    "pp_abi_version": (C.c_int, []),
    "pp_build_info": (C.c_char_p, []),
    "pp_last_error": (C.c_char_p, []),
    "pp_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pp_destroy": (C.c_int, [_vp]),
    "pp_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "pp_symbolic": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, C.c_int32, C.c_int64, _vp, _vp, _vp, _vp]),
    "pp_set_coupling_cliques": (C.c_int, [_vp, C.c_int32, _vp, _vp]),
    "pp_schur_size": (C.c_int64, [_vp]),
    "pp_coupling_stats": (C.c_int, [_vp, _i64p]),
    "pp_cplan_create": (C.c_int, [C.c_int32, C.c_int32, _vp, _vp, C.c_int64, _vp, _vp, C.c_int32, C.c_double, C.POINTER(_vp)]),
    "pp_cplan_get": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int64, _i64p]),
    "pp_cplan_destroy": (C.c_int, [_vp]),
    "pp_set_diagonal_classes": (C.c_int, [_vp, C.c_int64, _vp, C.c_int32, _vp]),
    "pp_set_shifts": (C.c_int, [_vp, _f64p]),
    "pp_value_uploads": (C.c_int64, [_vp]),
    "pp_numeric_local": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "pp_numeric_coupling": (C.c_int, [_vp, _vp, _vp]),
    "pp_inertia_local": (C.c_int, [_vp, _i64p]),
    "pp_inertia_coupling": (C.c_int, [_vp, _i64p]),
    "pp_solve_forward": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "pp_solve_backward": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "pp_residual_local": (C.c_int, [_vp, _vp, _vp]),
    "pp_residual_norms": (C.c_int, [_vp, _vp, _f64p, _vp]),
    "pp_refine_forward": (C.c_int, [_vp, _vp, _vp]),
    "pp_refine_backward": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "pp_schur_tail": (C.c_int, [_vp, _f64p]),
    "pp_host_copy": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "pp_host_equal": (C.c_int, [C.c_int64, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "pp_host_wake": (C.c_int, [C.c_int, C.c_int64]),
    "pp_peer_allreduce": (C.c_int, [C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_uint32, C.c_int64, _vp, _vp]),
    "pp_stage_values": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "pp_factor_bytes": (C.c_int64, [_vp]),
    "pp_local_dim": (C.c_int64, [_vp]),
    "pp_kernel_launches": (C.c_int64, [_vp]),
    "pp_profile": (C.c_int, [_vp, _f64p, _i64p, C.c_int]),
    "pp_plan_stats": (C.c_int, [_vp, C.c_int32, _i64p]),
    "pp_plan_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "pp_plan_set_ordering": (C.c_int, [C.c_int32]),
    "pp_plan_set_hint": (C.c_int, [_vp, C.c_int64]),
    "pp_plan_get": (C.c_int, [_vp, C.c_char_p, C.POINTER(_i32p), _i64p]),
    "pp_plan_scalar": (C.c_int, [_vp, C.c_char_p, _i64p]),
    "pp_plan_destroy": (C.c_int, [_vp]),
    "pp_debug_front": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, _i32p, _vp, _vp]),
    # interior-point vector kernels (N3)
    "pp_ipm_workspace_bytes": (C.c_int64, []),
    "pp_ipm_fill": (C.c_int, [_vp, C.c_int32, C.c_double, _vp]),
    "pp_ipm_fraction_to_boundary": (C.c_int, [C.c_int64, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_ipm_complementarity": (C.c_int, [C.c_int64, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_ipm_max_abs": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "pp_ipm_step": (C.c_int, [C.c_int64, _vp, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_ipm_axpy": (C.c_int, [C.c_int64, _vp, C.c_int32, _vp, _vp, _vp]),
}

_lib = None

try:   # pointer tables of many arrays in one call (csrc/fastptr.c, built by __graft_entry__.build())
    from . import _pp_fastptr as _fp
except ImportError:   # host bookkeeping only: the interpreter builds the same tables, slower
    _fp = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load ``libparapint_b200.so`` (once) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().pp_last_error().decode()


def np_ptr(arr):
    """void* of a C-contiguous numpy array (kept alive by the caller)."""
    return arr.ctypes.data_as(_vp)


PLAN_ARRAYS = ("rootcols", "col_ptr", "cols", "row_ptr", "rows", "rel", "parent", "nchild", "dcap", "dslot", "child_ptr",
               "child_idx", "root_children", "tiny_ptr", "tiny_idx", "med_ptr", "med_idx", "big_ptr", "big_idx", "ent_ptr",
               "tgt_row", "tgt_col", "tgt_src_ptr", "tgt_src", "root_row", "root_col", "root_src")
PLAN_SCALARS = ("n", "m", "nT", "DR", "ns", "nnz_l", "max_front", "l_total", "stack_cap", "nlevels")


class HostCopier:
    """Multi-threaded gather / scatter between many numpy arrays and one staging buffer (``pp_host_copy``).

    The pointer table is cached while the caller keeps handing in the *same* array objects (the usual case:
    an interior-point loop updates the KKT values in place); strong references are held so ids cannot be reused."""

    def __init__(self, threads=None):
        self.lib = load()
        if threads:
            self.threads = int(threads)
        else:
            # share the host cores with the other ranks of this node: this rank's share of the cores it may run on,
            # less one for the interpreter thread that drives the GPU
            self.threads = max(1, min(8, self._cores_per_rank() - 1))
        env = os.environ.get("PARAPINT_B200_WAKE_US")
        if env not in (None, ""):
            self.wake_us = float(env)
        else:
            # workers that spin through the GPU waits need cores of their own: with fewer than twelve per rank the pool
            # is woken by its jobs as before
            self.wake_us = 400.0 if self._cores_per_rank() >= 12 else 0.0
        self._keep = None
        self._tables = None
        self.stage = None   # (handle, stream getter): gather straight into a transfer (see copy)

    @staticmethod
    def _cores_per_rank():
        """Host cores this rank may count on (torchrun exports LOCAL_WORLD_SIZE)."""
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        try:
            cores = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            cores = os.cpu_count() or 2
        return cores // ranks

    def _table(self, arrays, offsets):
        import numpy as np
        keep = self._keep
        off = np.asarray(offsets, dtype=np.int64) * 8
        if keep is not None:
            same = _fp.same(arrays, keep) if _fp is not None else \
                (len(keep) == len(arrays) and all(a is b for a, b in zip(arrays, keep)))
            if same and np.array_equal(off, self._tables[1]):   # same arrays AND same destinations
                return self._tables
        n = len(arrays)
        ptr = np.empty(n, dtype=np.uintp)
        ln = np.empty(n, dtype=np.int64)
        if _fp is not None:
            if _fp.fill(arrays, ptr, ln, "d") != n:   # an entry is not a contiguous float64 buffer
                return None
        else:
            for k, a in enumerate(arrays):
                if a.dtype != np.float64 or not a.flags.c_contiguous:
                    return None
                ptr[k] = a.__array_interface__["data"][0]
                ln[k] = a.size * 8
        self._keep = list(arrays)
        self._tables = (ptr, off, ln, np_ptr(ptr), np_ptr(off), np_ptr(ln))
        return self._tables

    def wake(self, spin_us=None):
        """A copy follows shortly (after a wait for the GPU, or after walking the matrix): the pool's workers wake up
        now and spin for it, for at most ``spin_us`` microseconds (``pp_host_wake``; default: the
        ``PARAPINT_B200_WAKE_US`` environment variable, 400; 0 disables)."""
        us = self.wake_us if spin_us is None else spin_us
        if us > 0 and self.threads > 1:
            self.lib.pp_host_wake(self.threads, int(us * 1000))

    def all_equal(self, fresh, refs):
        """True when ``fresh[k]`` and ``refs[k]`` hold the same bytes for every k (index arrays of fresh leaves against
        the analysed pattern); None when a pair does not qualify for the byte comparison -- different types or
        lengths, not contiguous -- and the caller has to compare with numpy."""
        import numpy as np
        n = len(fresh)
        if n == 0:
            return True
        if n != len(refs):
            return None
        pa = np.empty(n, dtype=np.uintp)
        pb = np.empty(n, dtype=np.uintp)
        ln = np.empty(n, dtype=np.int64)
        if _fp is not None:
            if _fp.pairs(fresh, refs, pa, pb, ln) != n:
                return None
        else:
            for k, (a, b) in enumerate(zip(fresh, refs)):
                if a.dtype != b.dtype or a.size != b.size or not a.flags.c_contiguous or not b.flags.c_contiguous:
                    return None
                pa[k], pb[k], ln[k] = a.ctypes.data, b.ctypes.data, b.nbytes
        out = C.c_int(0)
        if self.lib.pp_host_equal(n, np_ptr(pa), np_ptr(pb), np_ptr(ln), self.threads, C.byref(out)) != 0:
            raise RuntimeError(f"pp_host_equal failed: {last_error()}")
        return bool(out.value)

    def copy(self, arrays, offsets, staging, to_staging=True):
        """``arrays[k]`` (float64, contiguous) <-> ``staging[offsets[k] : offsets[k] + arrays[k].size]``.
        Returns False when an array does not qualify (the caller falls back to numpy slicing)."""
        if not arrays:
            return True
        t = self._table(arrays, offsets)
        if t is None:
            return False
        if self.stage is not None and to_staging:
            # values of a factorisation: gather + host-to-device transfer pipelined by the library
            handle, stream = self.stage
            code = self.lib.pp_stage_values(handle, len(arrays), t[3], t[4], t[5], staging.ctypes.data, self.threads,
                                            4, stream())
            if code == 0:
                return True
            self.stage = None  # e.g. pattern with gaps: fall back to the plain gather below
        code = self.lib.pp_host_copy(len(arrays), t[3], t[4], t[5], staging.ctypes.data, 1 if to_staging else 0,
                                     self.threads)
        if code != 0:
            raise RuntimeError(f"pp_host_copy failed: {last_error()}")
        return True


def build_plan(n, m, rows, cols, fmax=-1, dmax=-1, min_sparse_n=-1, ordering=0, values=None):
    """Run the host symbolic analysis of one block and return its tables as a dict of numpy arrays.
    ``values``: representative values of the entries (the values hint of ``pp_symbolic``), or None."""
    import numpy as np
    lib = load()
    lib.pp_plan_set_ordering(ordering)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    if values is not None:
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.size != rows.size:
            raise ValueError("values must have one entry per (row, col) pair")
        if lib.pp_plan_set_hint(np_ptr(values), values.size) != 0:
            raise RuntimeError(last_error())
    else:
        lib.pp_plan_set_hint(None, 0)
    plan = _vp()
    if lib.pp_plan_create(n, m, rows.size, np_ptr(rows), np_ptr(cols), fmax, dmax, min_sparse_n, C.byref(plan)) != 0:
        raise RuntimeError(last_error())
    try:
        out = {}
        for name in PLAN_ARRAYS:
            ptr, ln = _i32p(), C.c_int64()
            if lib.pp_plan_get(plan, name.encode(), C.byref(ptr), C.byref(ln)) != 0:
                raise RuntimeError(last_error())
            out[name] = np.ctypeslib.as_array(ptr, shape=(ln.value,)).copy() if ln.value else np.zeros(0, dtype=np.int32)
        for name in PLAN_SCALARS:
            val = C.c_int64()
            if lib.pp_plan_scalar(plan, name.encode(), C.byref(val)) != 0:
                raise RuntimeError(last_error())
            out[name] = int(val.value)
        return out
    finally:
        lib.pp_plan_destroy(plan)


CPLAN_ARRAYS = ("scalars", "colptr", "rowidx", "block_n", "border_ptr", "border_rows", "dest_front", "dest_row",
                "dest_col", "perm_local", "perm_c")


def coupling_plan(m_c, clique_ptr, clique_rows, qrow, qcol, min_mc=-1, max_density=-1.0):
    """One level of the host analysis of a sparse coupling system (``pp_cplan_*``) as a dict of numpy arrays."""
    import numpy as np
    lib = load()
    cp = np.ascontiguousarray(clique_ptr, dtype=np.int64)
    cr = np.ascontiguousarray(clique_rows, dtype=np.int32)
    qr = np.ascontiguousarray(qrow, dtype=np.int32)
    qc = np.ascontiguousarray(qcol, dtype=np.int32)
    plan = _vp()
    if lib.pp_cplan_create(int(m_c), cp.size - 1, np_ptr(cp), np_ptr(cr), qr.size, np_ptr(qr), np_ptr(qc), int(min_mc),
                           float(max_density), C.byref(plan)) != 0:
        raise RuntimeError(last_error())
    try:
        out = {}
        for name in CPLAN_ARRAYS:
            ln = C.c_int64()
            if lib.pp_cplan_get(plan, name.encode(), None, 0, C.byref(ln)) != 0:
                raise RuntimeError(last_error())
            buf = np.zeros(max(ln.value, 1), dtype=np.int64)
            lib.pp_cplan_get(plan, name.encode(), np_ptr(buf), buf.size, C.byref(ln))
            out[name] = buf[: ln.value]
        sc = out.pop("scalars")
        out.update(sparse=bool(sc[0]), n_blocks=int(sc[1]), m_next=int(sc[2]), nnz=int(sc[3]), m_c=int(sc[4]))
        return out
    finally:
        lib.pp_cplan_destroy(plan)
