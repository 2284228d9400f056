/* py_fastq.c -- fastqdedup_b200._fastq: drop-in for the reference's fastqdedup._fastq
 * (src/fastqdedup/_fastqmodule.c:38-101, stub _fastq.pyi:17-21).  The arithmetic runs in
 * libfqd_b200's error_rate kernel; this file only converts arguments. */
#include "py_common.h"

static PyObject *
py_average_error_rate(PyObject *module, PyObject *args, PyObject *kwargs)
{
    static char *names[] = {"", "phred_offset", NULL};
    PyObject *scores = NULL;
    unsigned char offset = FQD_DEFAULT_PHRED_OFFSET;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O!|$b:average_error_rate", names,
                                     &PyUnicode_Type, &scores, &offset))
        return NULL;
    if (!PyUnicode_IS_COMPACT_ASCII(scores)) {
        PyErr_SetString(PyExc_ValueError, "phred_scores must be ASCII encoded.");
        return NULL;
    }
    fqd_context *ctx = fqd_py_context();
    if (ctx == NULL)
        return NULL;
    uint64_t offsets[2] = {0, (uint64_t)PyUnicode_GET_LENGTH(scores)};
    double result = 0.0;
    uint32_t bad_char = 0;
    int rc = fqd_average_error_rate(ctx, (const uint8_t *)PyUnicode_DATA(scores), offsets, 1,
                                    offset, &result, NULL, &bad_char);
    if (rc == FQD_ERR_PHRED) {
        /* built here rather than from the C string so that a NUL byte survives */
        PyErr_Format(PyExc_ValueError, "Character %c outside of valid phred range ('%c' to '%c')",
                     (int)bad_char, (int)offset, 126);
        return NULL;
    }
    if (rc != FQD_OK)
        return fqd_py_raise(rc);
    return PyFloat_FromDouble(result);
}

static PyMethodDef fastq_methods[] = {
    {"average_error_rate", (PyCFunction)(void (*)(void))py_average_error_rate,
     METH_VARARGS | METH_KEYWORDS,
     "average_error_rate($module, phred_scores, /, *, phred_offset=33)\n--\n\n"
     "Mean per-base error probability of an ASCII PHRED string (computed on the GPU)."},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef fastq_module = {PyModuleDef_HEAD_INIT, "_fastq", NULL, -1, fastq_methods};

PyMODINIT_FUNC
PyInit__fastq(void)
{
    PyObject *m = PyModule_Create(&fastq_module);
    if (m == NULL)
        return NULL;
    if (PyModule_AddIntConstant(m, "DEFAULT_PHRED_OFFSET", FQD_DEFAULT_PHRED_OFFSET) < 0)
        return NULL;
    return m;
}
