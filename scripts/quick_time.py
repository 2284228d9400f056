"""Scratch timing of the device-resident job at a few sizes (not the bench)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_device

ctx = _native.Context(0)
for name, n in [("cfg5", 2_000_000), ("cfg5", 10_000_000), ("cfg3", 4_000_000), ("cfg4", 2_000_000), ("cfg2", 5_000_000)]:
    cfg = synth.CONFIGS[name].scaled(n)
    t0 = time.time()
    keys, lens, quals = synth.SynthSource(cfg).reads()
    t1 = time.time()
    kp = ctx.upload(keys)
    qp = ctx.upload(quals) if quals is not None else None
    for it in range(3):
        st = cluster_device(ctx, n, kp, cfg.key_length, quals_ptr=qp, qual_length=cfg.key_length,
                            max_distance=cfg.max_distance, use_edit_distance=cfg.use_edit_distance,
                            method=cfg.method, max_average_error_rate=cfg.max_average_error_rate)
    d = st.as_dict()
    print(name, n, f"gen {t1-t0:.1f}s", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()}, flush=True)
    print(f"   -> {d['number_of_uniques']/d['ms_total']*1e3/1e6:.1f} M uniques/s", flush=True)
    ctx.device_free(kp)
    if qp: ctx.device_free(qp)
