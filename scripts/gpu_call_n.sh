#!/bin/bash
# GPU call: sharded bench at N GPUs (usage: gpu_call_n.sh N [tag])
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
FQD_WATCHDOG=600 FQD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
head -c 1200 gpurun_out/${TAG}_bench_n$N.json; grep "fqd trace" gpurun_out/${TAG}_bench_n$N.err | tail -3
