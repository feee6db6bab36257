/* _hostgather: packs a sequence of small NumPy matrices into one buffer.
 *
 * The reference hands its likelihood one dict {(parent, child): S x S ndarray} per rate category
 * (ML_gamma.pyx:22-36 indexes them edge by edge), 8 184 separate 2 x 2 arrays for a 1 024-taxon tree.
 * Before the single host -> device copy they have to become one contiguous block; bytes.join does that through the
 * buffer protocol at ~100 ns per array, this module reads the array structs directly (~10 ns per array).
 * Host-side glue only: no likelihood arithmetic happens here, and likelihood.py falls back to bytes.join with the
 * same bytes when the module has not been built.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <string.h>

/* one array into its place; 0 when it is not a plain float64 matrix of `each` elements */
static inline int put(PyObject* v, char* dst, npy_intp each, size_t bytes) {
  if (!PyArray_Check(v)) return 0;
  PyArrayObject* a = (PyArrayObject*)v;
  if (PyArray_TYPE(a) != NPY_DOUBLE || PyArray_SIZE(a) != each || !PyArray_IS_C_CONTIGUOUS(a) || !PyArray_ISALIGNED(a) ||
      !PyArray_ISNOTSWAPPED(a))
    return 0;
  memcpy(dst, PyArray_DATA(a), bytes);
  return 1;
}

/* pack(values, out) -> bool.  values: list or tuple; out: writable C-contiguous float64 ndarray whose size is
 * len(values) * (elements per matrix).  Returns False, leaving `out` unspecified, as soon as an element is not a
 * C-contiguous aligned float64 ndarray of exactly that many elements -- the caller then takes the general route. */
static PyObject* pack(PyObject* self, PyObject* args) {
  PyObject *values, *out_obj;
  if (!PyArg_ParseTuple(args, "OO", &values, &out_obj)) return NULL;
  if (!PyArray_Check(out_obj)) {
    PyErr_SetString(PyExc_TypeError, "out must be an ndarray");
    return NULL;
  }
  PyArrayObject* out = (PyArrayObject*)out_obj;
  if (PyArray_TYPE(out) != NPY_DOUBLE || !PyArray_IS_C_CONTIGUOUS(out) || !PyArray_ISWRITEABLE(out) || !PyArray_ISALIGNED(out)) {
    PyErr_SetString(PyExc_TypeError, "out must be a writable C-contiguous float64 ndarray");
    return NULL;
  }
  PyObject* seq = PySequence_Fast(values, "values must be a list or tuple");
  if (!seq) return NULL;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
  PyObject** items = PySequence_Fast_ITEMS(seq);
  const npy_intp total = PyArray_SIZE(out);
  if (n == 0 || total % n != 0) {
    Py_DECREF(seq);
    if (n == 0 && total == 0) Py_RETURN_TRUE;
    Py_RETURN_FALSE;
  }
  const npy_intp each = total / n;
  const size_t bytes = (size_t)each * sizeof(double);
  char* dst = (char*)PyArray_DATA(out);
  int ok = 1;
  for (Py_ssize_t i = 0; i < n; ++i) {
    if (!put(items[i], dst + (size_t)i * bytes, each, bytes)) { ok = 0; break; }
  }
  Py_DECREF(seq);
  if (ok) Py_RETURN_TRUE;
  Py_RETURN_FALSE;
}

/* pack_dict(table, keys, out) -> bool.  table: dict; keys: list.  True when the dict's keys, in its own order, are
 * exactly `keys` (same objects, or equal) and every value was copied into `out` in that order; False otherwise (`out`
 * unspecified).  One pass over the dict replaces list(table) == keys, list(table.values()) and pack(). */
static PyObject* pack_dict(PyObject* self, PyObject* args) {
  PyObject *table, *keys, *out_obj;
  if (!PyArg_ParseTuple(args, "OOO", &table, &keys, &out_obj)) return NULL;
  if (!PyDict_CheckExact(table) || !PyList_CheckExact(keys)) Py_RETURN_FALSE;
  if (!PyArray_Check(out_obj)) {
    PyErr_SetString(PyExc_TypeError, "out must be an ndarray");
    return NULL;
  }
  PyArrayObject* out = (PyArrayObject*)out_obj;
  if (PyArray_TYPE(out) != NPY_DOUBLE || !PyArray_IS_C_CONTIGUOUS(out) || !PyArray_ISWRITEABLE(out) || !PyArray_ISALIGNED(out)) {
    PyErr_SetString(PyExc_TypeError, "out must be a writable C-contiguous float64 ndarray");
    return NULL;
  }
  const Py_ssize_t n = PyList_GET_SIZE(keys);
  const npy_intp total = PyArray_SIZE(out);
  if (n == 0 || PyDict_GET_SIZE(table) != n || total % n != 0) Py_RETURN_FALSE;
  const npy_intp each = total / n;
  const size_t bytes = (size_t)each * sizeof(double);
  char* dst = (char*)PyArray_DATA(out);
  Py_ssize_t pos = 0, i = 0;
  PyObject *k, *v;
  int ok = 1;
  while (ok && PyDict_Next(table, &pos, &k, &v)) {
    if (i >= PyList_GET_SIZE(keys)) { ok = 0; break; }
    PyObject* want = PyList_GET_ITEM(keys, i);
    if (k != want) {
      /* equal but distinct key objects: compare (holding references -- a comparison may run arbitrary code) */
      Py_INCREF(k); Py_INCREF(v); Py_INCREF(want);
      const int eq = PyObject_RichCompareBool(k, want, Py_EQ);
      const int unchanged = PyDict_GET_SIZE(table) == n && PyList_GET_SIZE(keys) == n;
      if (eq < 0) PyErr_Clear();
      ok = eq > 0 && unchanged && put(v, dst + (size_t)i * bytes, each, bytes);
      Py_DECREF(k); Py_DECREF(v); Py_DECREF(want);
    } else {
      ok = put(v, dst + (size_t)i * bytes, each, bytes);
    }
    ++i;
  }
  if (ok && i == n) Py_RETURN_TRUE;
  Py_RETURN_FALSE;
}

static PyMethodDef methods[] = {
    {"pack_dict", pack_dict, METH_VARARGS,
     "pack_dict(table, keys, out) -> bool: copy a dict's float64 ndarrays into out if its key order is `keys`"},
    {"pack", pack, METH_VARARGS, "pack(values, out) -> bool: copy a sequence of equal-sized float64 ndarrays into out"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_hostgather", NULL, -1, methods};

PyMODINIT_FUNC PyInit__hostgather(void) {
  import_array();
  return PyModule_Create(&module);
}
