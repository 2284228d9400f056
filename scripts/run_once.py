"""One device-resident job (for ncu captures / traces): python scripts/run_once.py [n_reads] [iterations] [config]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from bench import generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.CONFIGS[sys.argv[3] if len(sys.argv) > 3 else "cfg5"].scaled(n)
L = cfg.key_length
keys = np.empty((n, L), dtype=np.uint8)
quals = np.empty((n, L), dtype=np.uint8) if cfg.quality_mix else None
generate_into(cfg, 0, n, keys, quals)
ctx = _native.Context(0)
kp = ctx.upload(keys)
qp = ctx.upload(quals) if quals is not None else None
del keys, quals
for _ in range(iters):
    st = cluster_device(ctx, n, kp, L, quals_ptr=qp, qual_length=L, max_distance=cfg.max_distance,
                        use_edit_distance=cfg.use_edit_distance, method=cfg.method,
                        max_average_error_rate=cfg.max_average_error_rate)
d = st.as_dict()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})
