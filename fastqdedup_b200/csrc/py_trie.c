/* py_trie.c -- fastqdedup_b200._trie: drop-in for the reference's fastqdedup._trie
 * (src/fastqdedup/_triemodule.c:596-1009, stub _trie.pyi:20-44).
 *
 * The Python-visible surface (constructor keyword, method names, argument formats,
 * exception types and messages the reference's tests assert on) is the reference's; the
 * object behind it is libfqd_b200's fqd_trie, whose neighbour search and clustering run on
 * the GPU.  This file only converts arguments and results.
 */
#include "py_common.h"

typedef struct {
    PyObject_HEAD
    fqd_trie *trie;
} TrieObject;

static void
Trie_dealloc(TrieObject *self)
{
    fqd_trie_free(self->trie);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

/* Trie(alphabet: str = "")   -- reference :613-642 */
static PyObject *
Trie_new(PyTypeObject *type, PyObject *args, PyObject *kwargs)
{
    static char *names[] = {"alphabet", NULL};
    PyObject *alphabet = NULL;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|O!:Trie.__new__", names,
                                     &PyUnicode_Type, &alphabet))
        return NULL;
    const uint8_t *chars = (const uint8_t *)"";
    size_t nchars = 0;
    if (alphabet != NULL) {
        if (!PyUnicode_IS_COMPACT_ASCII(alphabet)) {
            PyErr_SetString(PyExc_ValueError, "Alphabet should be an ASCII string.");
            return NULL;
        }
        chars = (const uint8_t *)PyUnicode_DATA(alphabet);
        nchars = (size_t)PyUnicode_GET_LENGTH(alphabet);
    }
    fqd_context *ctx = fqd_py_context();
    if (ctx == NULL)
        return NULL;
    TrieObject *self = (TrieObject *)type->tp_alloc(type, 0);
    if (self == NULL)
        return NULL;
    self->trie = NULL;
    int rc = fqd_trie_new(ctx, chars, nchars, &self->trie);
    if (rc != FQD_OK) {
        Py_DECREF(self);
        return fqd_py_raise(rc);
    }
    return (PyObject *)self;
}

/* add_sequence(sequence, /)  -- reference :677-706 */
static PyObject *
Trie_add_sequence(TrieObject *self, PyObject *sequence)
{
    if (!PyUnicode_CheckExact(sequence)) {
        PyErr_Format(PyExc_TypeError, "Sequence must be a str, got %s",
                     Py_TYPE(sequence)->tp_name);
        return NULL;
    }
    if (!PyUnicode_IS_COMPACT_ASCII(sequence)) {
        PyErr_SetString(PyExc_ValueError, "Sequence must consist only of ASCII characters");
        return NULL;
    }
    int rc = fqd_trie_add_sequence(self->trie, (const uint8_t *)PyUnicode_DATA(sequence),
                                   (size_t)PyUnicode_GET_LENGTH(sequence));
    if (rc != FQD_OK)
        return fqd_py_raise(rc);
    Py_RETURN_NONE;
}

/* contains_sequence(sequence, /, max_distance=0, use_edit_distance=False) -- :730-758 */
static PyObject *
Trie_contains_sequence(TrieObject *self, PyObject *args, PyObject *kwargs)
{
    static char *names[] = {"", "max_distance", "use_edit_distance", NULL};
    PyObject *sequence = NULL;
    int max_distance = 0, edit = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O!|ip:Trie.contains_sequence", names,
                                     &PyUnicode_Type, &sequence, &max_distance, &edit))
        return NULL;
    if (!PyUnicode_IS_COMPACT_ASCII(sequence)) {
        PyErr_SetString(PyExc_ValueError, "sequence must contain only ASCII characters");
        return NULL;
    }
    int32_t found = 0;
    int rc = fqd_trie_contains_sequence(self->trie, (const uint8_t *)PyUnicode_DATA(sequence),
                                        (size_t)PyUnicode_GET_LENGTH(sequence), max_distance,
                                        edit, &found);
    if (rc != FQD_OK)
        return fqd_py_raise(rc);
    return PyBool_FromLong(found);
}

/* pop_cluster(max_distance, use_edit_distance=False) -> list[tuple[int, str]] -- :778-897 */
static PyObject *
Trie_pop_cluster(TrieObject *self, PyObject *args, PyObject *kwargs)
{
    static char *names[] = {"max_distance", "use_edit_distance", NULL};
    int max_distance = 0, edit = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "i|p:Trie.pop_cluster", names,
                                     &max_distance, &edit))
        return NULL;
    uint64_t n = 0;
    int rc = fqd_trie_pop_cluster(self->trie, max_distance, edit, &n);
    if (rc != FQD_OK)
        return fqd_py_raise(rc);
    PyObject *cluster = PyList_New((Py_ssize_t)n);
    if (cluster == NULL)
        return NULL;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t count = 0;
        const uint8_t *seq = NULL;
        size_t len = 0;
        fqd_trie_cluster_item(self->trie, i, &count, &seq, &len);
        PyObject *str = PyUnicode_DecodeLatin1((const char *)seq, (Py_ssize_t)len, NULL);
        PyObject *cnt = PyLong_FromUnsignedLong(count);
        PyObject *item = (str && cnt) ? PyTuple_Pack(2, cnt, str) : NULL;
        Py_XDECREF(str);
        Py_XDECREF(cnt);
        if (item == NULL) {
            Py_DECREF(cluster);
            return NULL;
        }
        PyList_SET_ITEM(cluster, (Py_ssize_t)i, item);
    }
    return cluster;
}

/* memory_size() -> int -- :909-913 */
static PyObject *
Trie_memory_size(TrieObject *self, PyObject *Py_UNUSED(ignored))
{
    return PyLong_FromUnsignedLongLong(fqd_trie_memory_size(self->trie));
}

/* raw_stats() -> list[list[int]] -- :929-964 */
static PyObject *
Trie_raw_stats(TrieObject *self, PyObject *Py_UNUSED(ignored))
{
    size_t row_len = 0;
    size_t rows = fqd_trie_raw_stats(self->trie, NULL, 0, &row_len);
    uint64_t *buf = PyMem_Calloc((rows && row_len) ? rows * row_len : 1, sizeof(uint64_t));
    if (buf == NULL)
        return PyErr_NoMemory();
    fqd_trie_raw_stats(self->trie, buf, rows * row_len, &row_len);
    PyObject *result = PyList_New((Py_ssize_t)rows);
    for (size_t r = 0; result != NULL && r < rows; r++) {
        PyObject *row = PyList_New((Py_ssize_t)row_len);
        if (row == NULL) {
            Py_CLEAR(result);
            break;
        }
        for (size_t c = 0; c < row_len; c++)
            PyList_SET_ITEM(row, (Py_ssize_t)c, PyLong_FromUnsignedLongLong(buf[r * row_len + c]));
        PyList_SET_ITEM(result, (Py_ssize_t)r, row);
    }
    PyMem_Free(buf);
    return result;
}

static PyObject *
Trie_get_alphabet(TrieObject *self, void *closure)
{
    uint8_t buf[256];
    size_t n = fqd_trie_alphabet(self->trie, buf, sizeof buf);
    return PyUnicode_DecodeLatin1((const char *)buf, (Py_ssize_t)n, NULL);
}

static PyObject *
Trie_get_number_of_sequences(TrieObject *self, void *closure)
{
    return PyLong_FromUnsignedLongLong(fqd_trie_number_of_sequences(self->trie));
}

static PyGetSetDef Trie_getset[] = {
    {"alphabet", (getter)Trie_get_alphabet, NULL, "The alphabet this trie uses.", NULL},
    {"number_of_sequences", (getter)Trie_get_number_of_sequences, NULL,
     "The number of sequences stored in the trie.", NULL},
    {NULL}};

static PyMethodDef Trie_methods[] = {
    {"add_sequence", (PyCFunction)Trie_add_sequence, METH_O,
     "add_sequence($self, sequence, /)\n--\n\nAdds an ASCII sequence."},
    {"contains_sequence", (PyCFunction)(void (*)(void))Trie_contains_sequence,
     METH_VARARGS | METH_KEYWORDS,
     "contains_sequence($self, sequence, /, max_distance=0, use_edit_distance=False)\n--\n\n"
     "True when a stored sequence lies within max_distance of sequence."},
    {"pop_cluster", (PyCFunction)(void (*)(void))Trie_pop_cluster, METH_VARARGS | METH_KEYWORDS,
     "pop_cluster($self, max_distance, use_edit_distance=False)\n--\n\n"
     "Removes one cluster of sequences chained within max_distance and returns it as a\n"
     "list of (count, sequence) tuples."},
    {"memory_size", (PyCFunction)Trie_memory_size, METH_NOARGS,
     "memory_size($self)\n--\n\nBytes held by the staged sequences."},
    {"raw_stats", (PyCFunction)Trie_raw_stats, METH_NOARGS,
     "raw_stats($self)\n--\n\nPer-layer node statistics (list of lists)."},
    {NULL}};

static PyTypeObject Trie_Type = {
    PyVarObject_HEAD_INIT(NULL, 0)
    .tp_name = "_trie.Trie",
    .tp_basicsize = sizeof(TrieObject),
    .tp_dealloc = (destructor)Trie_dealloc,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_new = Trie_new,
    .tp_methods = Trie_methods,
    .tp_getset = Trie_getset,
};

static struct PyModuleDef trie_module = {PyModuleDef_HEAD_INIT, "_trie", NULL, -1, NULL};

PyMODINIT_FUNC
PyInit__trie(void)
{
    if (PyType_Ready(&Trie_Type) < 0)
        return NULL;
    PyObject *m = PyModule_Create(&trie_module);
    if (m == NULL)
        return NULL;
    Py_INCREF(&Trie_Type);
    if (PyModule_AddObject(m, "Trie", (PyObject *)&Trie_Type) < 0) {
        Py_DECREF(&Trie_Type);
        Py_DECREF(m);
        return NULL;
    }
    return m;
}
