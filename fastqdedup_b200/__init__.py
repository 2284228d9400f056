"""fastqdedup_b200 -- B200-native clustering hot path of fastqdedup behind the reference's
own extension-module API.  See DESIGN.md / INTEGRATION.md.

The native pieces (``libfqd_b200.so`` and the ``_trie`` / ``_distance`` / ``_fastq`` shims)
are built in-tree by ``python -m fastqdedup_b200.build``; without them every name below
raises ImportError on first use -- there is no Python or CPU fallback for the clustering
arithmetic.  Names resolve lazily (PEP 562), so importing the package, or one of its pure
Python helpers (``build``, ``synth``), loads no shared object: a process that only wants the
synthetic-workload generator (``bench.py --impl reference``) never maps the CUDA library.
"""
import importlib

# public name -> (submodule, attribute): the reference's module surface
# (src/fastqdedup/__init__.py:32-34 and its top-level functions) plus the batched job
_EXPORTS = {
    "within_distance": ("._distance", "within_distance"),
    "fastq_average_error_rate": ("._fastq", "average_error_rate"),
    "Trie": ("._trie", "Trie"),
    "ClusterResult": (".clustering", "ClusterResult"),
    "cluster_device": (".clustering", "cluster_device"),
    "cluster_keys": (".clustering", "cluster_keys"),
}
for _name in ("CLUSTER_DISSECTION_METHODS", "DEFAULT_CLUSTER_DISSECTION",
              "DEFAULT_MAX_AVERAGE_ERROR_RATE", "DEFAULT_MAX_DISTANCE", "DEFAULT_PREFIX",
              "argument_parser", "cluster_dissection_adjacency", "cluster_dissection_directional",
              "cluster_dissection_highest_count", "deduplicate_cluster", "initiate_logger",
              "length_string_to_slices", "main", "trie_stats"):
    _EXPORTS[_name] = (".frontend", _name)
del _name

__all__ = sorted(_EXPORTS)


def __getattr__(name):
    target = _EXPORTS.get(name)
    if target is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    try:
        value = getattr(importlib.import_module(target[0], __name__), target[1])
    except ImportError as missing:
        raise ImportError("fastqdedup_b200: the native library / extension modules are not built "
                          "(run `python -m fastqdedup_b200.build`); there is no CPU fallback") from missing
    globals()[name] = value
    return value


def __dir__():
    return sorted(set(globals()) | set(_EXPORTS))
