#!/bin/bash
# 2-GPU call: the new GPU tests, the 2-GPU NCCL test behind the C ABI, and the default bench under torchrun
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/smi2.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -k "any_image_size or rejects_images or cfg1_replay or two_gpu or second_gpu or block_integration" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_default_n2.json 2> gpurun_out/bench_default_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/bench_default_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5_n2.json 2> gpurun_out/bench_cfg5_n2.err; echo "bench cfg5 n2 rc=$?"
