#!/bin/bash
# 8 GPUs, bounded: the cfg5 legs alone (1.92 MB and 39.6 MB all-reduce through tsdr_chain_allreduce)
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-8}
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload cfg5_fullres --steps 3 --warmup 3 > gpurun_out/bench_cfg5_fullres_n$N.json 2> gpurun_out/bench_cfg5_fullres_n$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_cfg5_fullres_n$N.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_cfg5_n$N.err
