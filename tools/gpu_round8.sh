#!/bin/bash
# round 2, call 12: k_accumulate_rows (full-resolution accumulate through a bulk-copy ring) -- tests, launch list, bench leg
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests -m gpu -x -q -k "fullres or legacy_projection" > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
tail -4 gpurun_out/pytest_gpu8.log
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 16 --csv --log-file gpurun_out/launches8_cfg5_full.csv python tools/prof_chain.py cfg5 2 full > gpurun_out/launches8_cfg5_full.log 2>&1
timeout 200 python bench.py --workload cfg5_fullres --steps 2 --warmup 3 > gpurun_out/bench_cfg5_fullres_n1.json 2> gpurun_out/bench_cfg5_fullres_n1.err; echo "rc=$?"
