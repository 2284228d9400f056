"""Per-phase device times of the sharded plan under torchrun, for a list of library switches, on one
generation of the workload (measurement aid; bench.py is the number that counts).

    python -m torch.distributed.run --nproc-per-node N scripts/shard_probe.py [reads] [variant ...]

A variant is NAME or NAME=ENV1:VAL1,ENV2:VAL2.  Prints the library's FQD_TRACE lines of the last step of every
variant (every rank with FQD_TRACE=2) and the max-over-ranks device time.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench import PinnedArray, generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.multigpu import ShardComm, shard_bounds

VARIANTS = {
    "default": {},
    "ldg": {"FQD_PEER_LDG": "1"},
    "full_regions": {"FQD_SHARD_FULL_REGIONS": "1"},
    "replicated": {"FQD_SHARD_REPLICATED": "1"},
}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    args = sys.argv[1:]
    n = int(args[0]) if args and args[0].isdigit() else 100_000_000
    names = [a for a in args if not a.isdigit()] or ["default"]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synth.CONFIGS["cfg5"].scaled(n) if n != 100_000_000 else synth.CONFIGS["cfg5"]
    lib = _native.load()
    ctx = _native.Context(local)
    comm = ShardComm.from_torch_distributed(ctx, dist)
    b = shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    host = PinnedArray(lib, (hi - lo, cfg.key_length))
    generate_into(cfg, lo, hi, host.array, None, threads=max(2, 16 // world))
    d_keys = ctx.upload(host.array)
    d_bm = ctx.device_alloc(((hi - lo + 31) // 32 + 1) * 4)
    for name in names:
        env = dict(VARIANTS.get(name.split("=")[0], {}))
        if "=" in name:
            env.update(kv.split(":") for kv in name.split("=", 1)[1].split(","))
        for k in [k for k in os.environ if k.startswith("FQD_")]:
            os.environ.pop(k, None)
        os.environ.update(env)
        times = []
        for it in range(6):
            if it == 5:
                os.environ["FQD_TRACE"] = "2"
                dist.barrier()
                if rank == 0:
                    print(f"==== variant {name} {env}", file=sys.stderr, flush=True)
                dist.barrier()
            st = comm.cluster_device(hi - lo, lo, d_keys, cfg.key_length, max_distance=cfg.max_distance, method=cfg.method,
                                     bitmap_ptr=d_bm)
            t = torch.tensor([st.ms_total], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= 2:
                times.append(float(t.item()))
        os.environ.pop("FQD_TRACE", None)
        dist.barrier()
        time.sleep(0.2)
        if rank == 0:
            print(f"==== variant {name}: max-over-ranks ms/step {np.mean(times[:3]):.3f} (traced step {times[3]:.3f}), "
                  f"U {st.number_of_uniques} selected {st.number_selected} plan_flags {st.plan_flags}", file=sys.stderr, flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
