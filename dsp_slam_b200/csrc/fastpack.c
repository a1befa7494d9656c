/* _fastpack: marshal a list of detection dicts into a DspgnObjectIn array without per-field Python work.
 *
 * The reference's boundary is native too: pybind11's Eigen casters turn Eigen matrices into numpy arrays before
 * Optimizer.reconstruct_object runs (pybind11/eigen.h).  Here the opposite direction -- numpy arrays to the C ABI's
 * pointer + stride records (include/dspgn.h: DspgnObjectIn) -- is a CPython extension function, because the pure-Python
 * version (dsp_slam_b200/optimizer.py: BatchSolver._pack, kept as the general fallback) costs ~4 us per object.
 *
 *   pack(objs: list[dict], code_len: int, out_addr: int) -> bool
 * Fills out[0..len(objs)) (zero-initialised by the caller).  Returns False -- and the caller falls back to the Python
 * path -- whenever an object needs anything beyond pointer extraction: non-float32 data, missing buffer protocol, short
 * codes, non-contiguous 1-D arrays, unexpected shapes.  No data is copied and no reference is kept: the arrays stay
 * alive through `objs`, which the caller holds for the duration of the library call.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>
#include "dspgn.h"

static PyObject *k_T, *k_pts, *k_rays, *k_depth, *k_code, *k_scale, *k_class, *k_pixels, *k_invk, *k_tcw;

/* float32 buffer with `ndim` dimensions: data pointer, shape and element strides; 0 on success, 1 = not eligible */
static int f32_view(PyObject* o, int ndim, const float** data, Py_ssize_t* shape, Py_ssize_t* estride) {
  Py_buffer v;
  if (!PyObject_CheckBuffer(o)) return 1;
  if (PyObject_GetBuffer(o, &v, PyBUF_STRIDES | PyBUF_FORMAT) != 0) { PyErr_Clear(); return 1; }
  int bad = v.ndim != ndim || v.itemsize != 4 || v.format == NULL ||
            !((v.format[0] == 'f' && v.format[1] == 0) || ((v.format[0] == '<' || v.format[0] == '=') && v.format[1] == 'f' && v.format[2] == 0));
  if (!bad) {
    *data = (const float*)v.buf;
    for (int d = 0; d < ndim; ++d) {
      shape[d] = v.shape[d];
      if (v.strides[d] % 4) { bad = 1; break; }
      estride[d] = v.strides[d] / 4;
    }
  }
  PyBuffer_Release(&v);
  return bad;
}

static PyObject* fp_pack(PyObject* self, PyObject* args) {
  PyObject* objs; long code_len; unsigned long long addr;
  (void)self;
  if (!PyArg_ParseTuple(args, "OlK", &objs, &code_len, &addr)) return NULL;
  if (!PyList_Check(objs)) Py_RETURN_FALSE;
  DspgnObjectIn* out = (DspgnObjectIn*)(uintptr_t)addr;
  const Py_ssize_t n = PyList_GET_SIZE(objs);
  Py_ssize_t sh[2], st[2];
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* o = PyList_GET_ITEM(objs, i);
    if (!PyDict_Check(o)) Py_RETURN_FALSE;
    DspgnObjectIn* e = out + i;
    PyObject* v;
    const float* p;
    e->scale = 1.0f;
    if (!(v = PyDict_GetItem(o, k_T)) || f32_view(v, 2, &p, sh, st) || sh[0] != 4 || sh[1] != 4) Py_RETURN_FALSE;
    e->t_cam_obj = p; e->t_rs = (int32_t)st[0]; e->t_cs = (int32_t)st[1];
    if (!(v = PyDict_GetItem(o, k_pts)) || f32_view(v, 2, &p, sh, st) || sh[1] != 3 || sh[0] > 0x7fffffff) Py_RETURN_FALSE;
    e->pts = p; e->n_pts = (int32_t)sh[0]; e->pts_rs = (int32_t)st[0]; e->pts_cs = (int32_t)st[1];
    PyObject* px = PyDict_GetItem(o, k_pixels);
    int have_rays = 0;
    if (px && px != Py_None) {
      if (f32_view(px, 2, &p, sh, st) || sh[1] != 2) Py_RETURN_FALSE;
      if (sh[0] > 0) {
        Py_ssize_t s2[2], t2[2]; const float* kp;
        e->pixels = p; e->n_rays = (int32_t)sh[0]; e->pix_rs = (int32_t)st[0]; e->pix_cs = (int32_t)st[1];
        if (!(v = PyDict_GetItem(o, k_invk)) || f32_view(v, 2, &kp, s2, t2) || s2[0] != 3 || s2[1] != 3 || t2[0] != 3 || t2[1] != 1) Py_RETURN_FALSE;
        e->inv_k = kp;
        have_rays = 1;
      }
    }
    if (!have_rays && (v = PyDict_GetItem(o, k_rays)) && v != Py_None) {
      if (f32_view(v, 2, &p, sh, st) || (sh[0] > 0 && sh[1] != 3)) Py_RETURN_FALSE;
      if (sh[0] > 0) { e->rays = p; e->n_rays = (int32_t)sh[0]; e->rays_rs = (int32_t)st[0]; e->rays_cs = (int32_t)st[1]; have_rays = 1; }
    }
    if (have_rays) {
      v = PyDict_GetItem(o, k_depth);
      if (v && v != Py_None) {
        if (f32_view(v, 1, &p, sh, st) || (sh[0] > 1 && st[0] != 1)) Py_RETURN_FALSE;
        if (sh[0] > 0) { e->depth = p; e->n_depth = (int32_t)sh[0]; }
      }
    }
    if ((v = PyDict_GetItem(o, k_code)) && v != Py_None) {
      if (f32_view(v, 1, &p, sh, st) || sh[0] < code_len || (sh[0] > 1 && st[0] != 1)) Py_RETURN_FALSE;
      e->code = p;
    }
    if ((v = PyDict_GetItem(o, k_tcw)) && v != Py_None) {
      if (f32_view(v, 2, &p, sh, st) || sh[0] != 4 || sh[1] != 4 || st[0] != 4 || st[1] != 1) Py_RETURN_FALSE;
      e->t_cam_world = p;
    }
    if ((v = PyDict_GetItem(o, k_scale)) && v != Py_None) {
      const double s = PyFloat_AsDouble(v);
      if (s == -1.0 && PyErr_Occurred()) { PyErr_Clear(); Py_RETURN_FALSE; }
      e->scale = (float)s;
    }
    if ((v = PyDict_GetItem(o, k_class)) && v != Py_None) {
      const long c = PyLong_AsLong(v);
      if (c == -1 && PyErr_Occurred()) { PyErr_Clear(); Py_RETURN_FALSE; }
      e->class_id = (int32_t)c;
    }
  }
  Py_RETURN_TRUE;
}

static PyMethodDef methods[] = {{"pack", fp_pack, METH_VARARGS, "pack(objs, code_len, out_addr) -> bool"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "DspgnObjectIn marshalling", -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__fastpack(void) {
  k_T = PyUnicode_InternFromString("t_cam_obj"); k_pts = PyUnicode_InternFromString("pts");
  k_rays = PyUnicode_InternFromString("rays"); k_depth = PyUnicode_InternFromString("depth");
  k_code = PyUnicode_InternFromString("code"); k_scale = PyUnicode_InternFromString("scale");
  k_class = PyUnicode_InternFromString("class_id"); k_pixels = PyUnicode_InternFromString("pixels");
  k_invk = PyUnicode_InternFromString("inv_k"); k_tcw = PyUnicode_InternFromString("t_cam_world");
  if (sizeof(DspgnObjectIn) == 0) return NULL;
  return PyModule_Create(&moddef);
}
