#!/bin/bash
# 2-GPU validation: default bench line (also.cfg5 / cfg5_fullres / cfg4 with collectives), cfg5 with 125-frame blocks, 2-GPU tests
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/bench_default_n2.json 2> gpurun_out/bench_default_n2.err; echo "rc=$?"
tail -3 gpurun_out/bench_default_n2.err
TSDR_BENCH_CFG5_FRAMES=250 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_250_n2.json 2> gpurun_out/bench_cfg5_250_n2.err; echo "rc=$?"
timeout 200 python -m pytest tests -m gpu -x -q -k "two_gpu or second_gpu" > gpurun_out/pytest_2gpu.log 2>&1; tail -3 gpurun_out/pytest_2gpu.log
