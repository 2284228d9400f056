"""fastqdedup_b200 -- B200-native clustering hot path of fastqdedup behind the reference's
own extension-module API.  See DESIGN.md / INTEGRATION.md.

The native pieces (``libfqd_b200.so`` and the ``_trie`` / ``_distance`` / ``_fastq`` shims)
are built in-tree by ``python -m fastqdedup_b200.build``; without them every name of this
package raises ImportError -- there is no Python or CPU fallback for the clustering arithmetic
(only the ``build`` submodule is usable, which is how a clean tree gets built).
"""
try:
    from ._distance import within_distance            # noqa: F401  (reference __init__.py:32)
    from ._fastq import average_error_rate as fastq_average_error_rate  # noqa: F401  (:33)
    from ._trie import Trie                            # noqa: F401  (:34)
    from .clustering import ClusterResult, cluster_device, cluster_keys  # noqa: F401
    from .frontend import (                            # noqa: F401
        CLUSTER_DISSECTION_METHODS,
        DEFAULT_CLUSTER_DISSECTION,
        DEFAULT_MAX_AVERAGE_ERROR_RATE,
        DEFAULT_MAX_DISTANCE,
        DEFAULT_PREFIX,
        argument_parser,
        cluster_dissection_adjacency,
        cluster_dissection_directional,
        cluster_dissection_highest_count,
        deduplicate_cluster,
        initiate_logger,
        length_string_to_slices,
        main,
        trie_stats,
    )
except ImportError as _missing:   # native pieces not built (yet)
    _missing_native = _missing

    def __getattr__(name):
        import importlib.util
        if not name.startswith("_") and importlib.util.find_spec(f"{__name__}.{name}") is not None:
            raise AttributeError(name)   # a pure-Python submodule (build, synth ...): the import system loads it
        raise ImportError("fastqdedup_b200: the native library / extension modules are not built "
                          "(run `python -m fastqdedup_b200.build`); there is no CPU fallback") from _missing_native
