"""One device-resident config-5 job (for ncu captures): python scripts/run_once.py [n_reads] [iterations]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from bench import generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.CONFIGS["cfg5"].scaled(n)
L = cfg.key_length
keys = np.empty((n, L), dtype=np.uint8)
generate_into(cfg, 0, n, keys)
ctx = _native.Context(0)
kp = ctx.upload(keys)
del keys
for _ in range(iters):
    st = cluster_device(ctx, n, kp, L, max_distance=1, use_edit_distance=False, method="directional", max_average_error_rate=1.0)
d = st.as_dict()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})
