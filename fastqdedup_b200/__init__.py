"""fastqdedup_b200 -- B200-native clustering hot path of fastqdedup behind the reference's
own extension-module API.  See DESIGN.md / INTEGRATION.md.

The native pieces (``libfqd_b200.so`` and the ``_trie`` / ``_distance`` / ``_fastq`` shims)
are built in-tree by ``python -m fastqdedup_b200.build``; importing this package without
them fails loudly -- there is no Python or CPU fallback for the clustering arithmetic.
"""
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
