#!/bin/bash
# multi-GPU evidence: the default bench line (carries also.cfg5 / cfg5_fullres / cfg4) and the cfg5 leg alone, N = $1
N=${1:-8}
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err; echo "rc=$?"
tail -3 gpurun_out/bench_default_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "rc=$?"
tail -3 gpurun_out/bench_cfg5_n$N.err
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu or second_gpu" > gpurun_out/pytest_2gpu.log 2>&1; tail -3 gpurun_out/pytest_2gpu.log; fi
