"""Device-resident timing of one config under several environment-switch settings, inputs generated once:
python scripts/variants.py <config> <n_reads> "<VAR=val,VAR2=val2>" "<...>" ...   ("-" = defaults)
Prints per setting the stage times (CUDA events inside the library) and the result counters, which must not change."""
import os, sys, hashlib
sys.path.insert(0, ".")
import numpy as np
from bench import generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_device

name, n = sys.argv[1], int(sys.argv[2])
settings = sys.argv[3:] or ["-"]
cfg = synth.CONFIGS[name].scaled(n)
L = cfg.key_length
keys = np.empty((n, L), dtype=np.uint8)
quals = np.empty((n, L), dtype=np.uint8) if cfg.quality_mix else None
generate_into(cfg, 0, n, keys, quals)
ctx = _native.Context(0)
kp = ctx.upload(keys)
qp = ctx.upload(quals) if quals is not None else None
words = (n + 31) // 32
bm = ctx.device_alloc(words * 4)
del keys, quals
ref = None
for setting in settings:
    added, dist = [], cfg.max_distance
    if setting != "-":
        for kv in setting.split(","):
            k, v = kv.split("=")
            if k == "D":            # pseudo-variable: the job's max distance
                dist = int(v)
                continue
            os.environ[k] = v
            added.append(k)
    runs = []
    for it in range(5):
        st = cluster_device(ctx, n, kp, L, quals_ptr=qp, qual_length=L, max_distance=dist,
                            use_edit_distance=cfg.use_edit_distance, method=cfg.method,
                            max_average_error_rate=cfg.max_average_error_rate, bitmap_ptr=bm)
        runs.append(st.as_dict())
    for k in added:
        del os.environ[k]
    d = runs[-1]
    digest = hashlib.sha256(ctx.download(bm, words * 4, np.uint32).tobytes()).hexdigest()[:16]
    res = (d["number_of_uniques"], d["number_of_clusters"], d["number_selected"], digest)
    ref = ref if (ref and ref[0] == dist) else (dist, res)
    same = ref[1] == res
    ms = {k: round(float(np.mean([r[k] for r in runs[2:]])), 3) for k in d if k.startswith("ms_")}
    print(f"{setting:40s} {'OK ' if same else 'RESULT DIFFERS '}{res} {ms}", flush=True)
