/* py_common.h -- shared by the three CPython shims: status code -> exception mapping and
 * access to the library's process-wide GPU context. */
#ifndef FQD_PY_COMMON_H
#define FQD_PY_COMMON_H
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include "fqd_b200.h"

/* Raise the Python exception the reference would raise for this failure class. */
static PyObject *
fqd_py_raise(int rc)
{
    PyObject *type;
    switch (rc) {
    case FQD_ERR_ARG:
    case FQD_ERR_PHRED:
        type = PyExc_ValueError;
        break;
    case FQD_ERR_NOMEM:
        type = PyExc_MemoryError;
        break;
    case FQD_ERR_LOOKUP:
        type = PyExc_LookupError;
        break;
    case FQD_ERR_UNSUPPORTED:
        type = PyExc_NotImplementedError;
        break;
    default:
        type = PyExc_RuntimeError;
    }
    PyErr_SetString(type, fqd_last_error());
    return NULL;
}

/* The GPU context; NULL with an exception set when no B200 is usable (no CPU fallback). */
static fqd_context *
fqd_py_context(void)
{
    fqd_context *ctx = NULL;
    int rc = fqd_default_context(&ctx);
    if (rc != FQD_OK) {
        fqd_py_raise(rc);
        return NULL;
    }
    return ctx;
}

#endif
