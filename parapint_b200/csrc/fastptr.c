/* Pointer tables of many numpy arrays in one call (host binding helper, CPython C API).
 *
 * The plugin hands the C ABI (include/parapint_b200.h: pp_host_copy, pp_host_equal, pp_stage_values) tables of
 * host addresses: one entry per COO leaf of the KKT matrix.  parapint builds new leaf objects every iteration
 * (reference parapint/interfaces/interface.py:432-491), so the tables are rebuilt every iteration as well, and
 * `ndarray.ctypes.data` costs 0.5-2 us per array in the interpreter -- more than the copy it prepares.  Here the
 * buffer protocol yields address and length at ~40 ns per array.
 *
 * Only addresses and lengths are produced; no arithmetic of the solver lives here.  If this module is missing the
 * host code uses the interpreter path (same tables, slower).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

#include <string.h>

static int same_format(const char *a, const char *b) {
  if (!a) a = "B";
  if (!b) b = "B";
  return strcmp(a, b) == 0;
}

/* fill(seq, ptr_out, len_out, fmt) -> number of entries written, or -1 when an entry does not qualify.
 *   seq      sequence of objects exporting a C-contiguous buffer
 *   ptr_out  writable buffer of >= len(seq) uintptr_t : start addresses
 *   len_out  writable buffer of >= len(seq) int64_t   : lengths in bytes
 *   fmt      required struct format of every entry ("d" = float64; "" = any); an entry of another type, a
 *            non-contiguous one or one without the buffer protocol makes the call return -1 (the caller then takes
 *            its general path)
 */
static PyObject *fill(PyObject *self, PyObject *args) {
  PyObject *seq, *ptr_obj, *len_obj;
  const char *fmt = "";
  if (!PyArg_ParseTuple(args, "OOO|s", &seq, &ptr_obj, &len_obj, &fmt)) return NULL;
  PyObject *fast = PySequence_Fast(seq, "fill: first argument must be a sequence");
  if (!fast) return NULL;
  Py_buffer pb, lb;
  if (PyObject_GetBuffer(ptr_obj, &pb, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) {
    Py_DECREF(fast);
    return NULL;
  }
  if (PyObject_GetBuffer(len_obj, &lb, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) {
    PyBuffer_Release(&pb);
    Py_DECREF(fast);
    return NULL;
  }
  Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  long long written = -1;
  if ((Py_ssize_t)(pb.len / sizeof(uintptr_t)) < n || (Py_ssize_t)(lb.len / sizeof(int64_t)) < n) {
    PyErr_SetString(PyExc_ValueError, "fill: output tables are too small");
  } else {
    uintptr_t *ptr = (uintptr_t *)pb.buf;
    int64_t *len = (int64_t *)lb.buf;
    PyObject **items = PySequence_Fast_ITEMS(fast);
    Py_ssize_t k = 0;
    for (; k < n; ++k) {
      Py_buffer v;
      if (PyObject_GetBuffer(items[k], &v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
        PyErr_Clear();
        break;
      }
      int ok = fmt[0] == 0 || same_format(v.format, fmt);
      ptr[k] = (uintptr_t)v.buf;
      len[k] = (int64_t)v.len;
      PyBuffer_Release(&v);
      if (!ok) break;
    }
    written = k == n ? (long long)n : -1;
  }
  PyBuffer_Release(&lb);
  PyBuffer_Release(&pb);
  Py_DECREF(fast);
  if (PyErr_Occurred()) return NULL;
  return PyLong_FromLongLong(written);
}

/* pairs(seq_a, seq_b, pa_out, pb_out, len_out) -> number of pairs written, or -1 when a pair does not qualify:
 * entry k of both sequences must be C-contiguous buffers of the same type and the same length (the byte comparison
 * of pp_host_equal is only meaningful then). */
static PyObject *pairs(PyObject *self, PyObject *args) {
  PyObject *a, *b, *pa_obj, *pb_obj, *len_obj;
  if (!PyArg_ParseTuple(args, "OOOOO", &a, &b, &pa_obj, &pb_obj, &len_obj)) return NULL;
  PyObject *fa = PySequence_Fast(a, "pairs: arguments must be sequences");
  if (!fa) return NULL;
  PyObject *fb = PySequence_Fast(b, "pairs: arguments must be sequences");
  if (!fb) {
    Py_DECREF(fa);
    return NULL;
  }
  Py_buffer o[3];
  PyObject *outs[3] = {pa_obj, pb_obj, len_obj};
  int got = 0;
  for (; got < 3; ++got)
    if (PyObject_GetBuffer(outs[got], &o[got], PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) break;
  long long written = -1;
  if (got == 3) {
    Py_ssize_t n = PySequence_Fast_GET_SIZE(fa);
    if (n != PySequence_Fast_GET_SIZE(fb)) {
      written = -1;
    } else if ((Py_ssize_t)(o[0].len / sizeof(uintptr_t)) < n || (Py_ssize_t)(o[1].len / sizeof(uintptr_t)) < n ||
               (Py_ssize_t)(o[2].len / sizeof(int64_t)) < n) {
      PyErr_SetString(PyExc_ValueError, "pairs: output tables are too small");
    } else {
      uintptr_t *pa = (uintptr_t *)o[0].buf, *pb = (uintptr_t *)o[1].buf;
      int64_t *len = (int64_t *)o[2].buf;
      PyObject **ia = PySequence_Fast_ITEMS(fa), **ib = PySequence_Fast_ITEMS(fb);
      Py_ssize_t k = 0;
      for (; k < n; ++k) {
        Py_buffer va, vb;
        if (PyObject_GetBuffer(ia[k], &va, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
          PyErr_Clear();
          break;
        }
        if (PyObject_GetBuffer(ib[k], &vb, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
          PyErr_Clear();
          PyBuffer_Release(&va);
          break;
        }
        int ok = va.len == vb.len && same_format(va.format, vb.format);
        pa[k] = (uintptr_t)va.buf;
        pb[k] = (uintptr_t)vb.buf;
        len[k] = (int64_t)va.len;
        PyBuffer_Release(&va);
        PyBuffer_Release(&vb);
        if (!ok) break;
      }
      written = k == n ? (long long)n : -1;
    }
  }
  for (int k = 0; k < got; ++k) PyBuffer_Release(&o[k]);
  Py_DECREF(fa);
  Py_DECREF(fb);
  if (got < 3 || PyErr_Occurred()) return NULL;
  return PyLong_FromLongLong(written);
}

/* same(seq_a, seq_b) -> True when both sequences have the same length and hold the same objects (identity). */
static PyObject *same(PyObject *self, PyObject *args) {
  PyObject *a, *b;
  if (!PyArg_ParseTuple(args, "OO", &a, &b)) return NULL;
  PyObject *fa = PySequence_Fast(a, "same: arguments must be sequences");
  if (!fa) return NULL;
  PyObject *fb = PySequence_Fast(b, "same: arguments must be sequences");
  if (!fb) {
    Py_DECREF(fa);
    return NULL;
  }
  Py_ssize_t n = PySequence_Fast_GET_SIZE(fa);
  int eq = n == PySequence_Fast_GET_SIZE(fb);
  if (eq) {
    PyObject **ia = PySequence_Fast_ITEMS(fa), **ib = PySequence_Fast_ITEMS(fb);
    for (Py_ssize_t k = 0; k < n; ++k)
      if (ia[k] != ib[k]) {
        eq = 0;
        break;
      }
  }
  Py_DECREF(fa);
  Py_DECREF(fb);
  if (eq) Py_RETURN_TRUE;
  Py_RETURN_FALSE;
}

static PyMethodDef methods[] = {
    {"fill", fill, METH_VARARGS, "fill(seq, ptr_out, len_out, fmt=\"\"): addresses and byte lengths of the buffers in seq (fmt: required struct format, e.g. \"d\")"},
    {"pairs", pairs, METH_VARARGS, "pairs(a, b, pa_out, pb_out, len_out): address tables of two sequences of equally typed, equally long buffers"},
    {"same", same, METH_VARARGS, "same(a, b): both sequences hold the same objects"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pp_fastptr", "pointer tables of numpy arrays", -1, methods};

PyMODINIT_FUNC PyInit__pp_fastptr(void) { return PyModule_Create(&moduledef); }
