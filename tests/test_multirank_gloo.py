"""world_size-2 run of the host-side logic of the multi-GPU path on CPU (gloo): shard
bounds, hand-over of the 128-byte communicator id from rank 0, gathering of the per-rank keep
bitmaps into the global record order.  The device work itself is covered by
tests/test_gpu_sharded.py (virtual ranks on one GPU) and by bench.py --gpus N (NCCL)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from fastqdedup_b200.multigpu import shard_bounds, unpack_bitmap

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
n = 1003
bounds = shard_bounds(n, world)
lo, hi = bounds[rank], bounds[rank + 1]

# (1) the communicator id travels from rank 0 exactly like ShardComm.from_torch_distributed does it
box = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
assert box[0] == bytes(range(128))

# (2) every rank owns a contiguous slice; together they tile [0, n)
sizes = [None] * world
dist.all_gather_object(sizes, (lo, hi))
assert sizes[0][0] == 0 and sizes[-1][1] == n and all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))

# (3) per-rank bitmaps (bit t%32 of word t/32 over LOCAL indices) stitch into the global mask
rng = np.random.default_rng(7)
truth = rng.random(n) < 0.3
local = truth[lo:hi]
words = np.packbits(local, bitorder="little")
words = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.uint32)
gathered = [None] * world
dist.all_gather_object(gathered, (words, hi - lo))
mask = np.concatenate([unpack_bitmap(w, m) for w, m in gathered])
assert np.array_equal(mask, truth)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_host_logic(tmp_path):
    import socket
    import subprocess
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in out, out
