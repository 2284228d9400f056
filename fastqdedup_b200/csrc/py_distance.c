/* py_distance.c -- fastqdedup_b200._distance: drop-in for fastqdedup._distance
 * (src/fastqdedup/_distancemodule.c:46-118, stub _distance.pyi:19-21).  The predicate is
 * evaluated by libfqd_b200's within_distance kernel. */
#include "py_common.h"

static PyObject *
py_within_distance(PyObject *module, PyObject *args, PyObject *kwargs)
{
    static char *names[] = {"", "", "max_distance", "use_edit_distance", NULL};
    PyObject *s1 = NULL, *s2 = NULL;
    int max_distance = 0, edit = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O!O!i|p:within_distance", names,
                                     &PyUnicode_Type, &s1, &PyUnicode_Type, &s2,
                                     &max_distance, &edit))
        return NULL;
    if (PyUnicode_KIND(s1) != PyUnicode_1BYTE_KIND) {
        PyErr_SetString(PyExc_ValueError, "string1 must be ASCII or latin-1 encoded.");
        return NULL;
    }
    if (PyUnicode_KIND(s2) != PyUnicode_1BYTE_KIND) {
        PyErr_SetString(PyExc_ValueError, "string2 must be ASCII or latin-1 encoded.");
        return NULL;
    }
    fqd_context *ctx = fqd_py_context();
    if (ctx == NULL)
        return NULL;
    uint64_t off1[2] = {0, (uint64_t)PyUnicode_GET_LENGTH(s1)};
    uint64_t off2[2] = {0, (uint64_t)PyUnicode_GET_LENGTH(s2)};
    uint8_t out = 0;
    int rc = fqd_within_distance(ctx, PyUnicode_1BYTE_DATA(s1), off1, PyUnicode_1BYTE_DATA(s2),
                                 off2, 1, max_distance, edit, &out);
    if (rc != FQD_OK)
        return fqd_py_raise(rc);
    return PyBool_FromLong(out);
}

static PyMethodDef distance_methods[] = {
    {"within_distance", (PyCFunction)(void (*)(void))py_within_distance,
     METH_VARARGS | METH_KEYWORDS,
     "within_distance($module, string1, string2, /, max_distance, use_edit_distance=False)\n--\n\n"
     "True when the Hamming (or, with use_edit_distance, Levenshtein) distance of the two\n"
     "strings is at most max_distance.  Strings of unequal length are never within Hamming\n"
     "distance.  Evaluated on the GPU."},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef distance_module = {PyModuleDef_HEAD_INIT, "_distance", NULL, -1,
                                             distance_methods};

PyMODINIT_FUNC
PyInit__distance(void)
{
    return PyModule_Create(&distance_module);
}
