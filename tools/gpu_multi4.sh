#!/bin/bash
# 4 GPUs: the default bench line exactly as the driver's scaling run launches it (all legs, watchdog armed)
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_default_n4.json 2> gpurun_out/bench_default_n4.err; echo "rc=$?"
tail -2 gpurun_out/bench_default_n4.err
