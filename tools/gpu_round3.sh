#!/bin/bash
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -30 gpurun_out/pytest_gpu3.log
timeout 900 python bench.py > gpurun_out/bench_default3.json 2> gpurun_out/bench_default3.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default3.err
