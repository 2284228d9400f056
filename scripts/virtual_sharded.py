"""The tile-sharded plan as virtual ranks on ONE GPU (for ncu captures of its kernels):
python scripts/virtual_sharded.py <config> <n_reads> <world> [iterations]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from bench import generate_into
from fastqdedup_b200 import synth
from fastqdedup_b200.multigpu import cluster_keys_sharded_local

name, n, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
cfg = synth.CONFIGS[name].scaled(n)
L = cfg.key_length
keys = np.empty((n, L), dtype=np.uint8)
quals = np.empty((n, L), dtype=np.uint8) if cfg.quality_mix else None
generate_into(cfg, 0, n, keys, quals)
for _ in range(iters):
    res = cluster_keys_sharded_local(keys, quals, cfg.max_distance, cfg.use_edit_distance, cfg.method,
                                     cfg.max_average_error_rate, world=world, want_uniques=False)
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.stats.items()})
