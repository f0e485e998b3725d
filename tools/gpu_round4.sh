#!/bin/bash
# round 2, call 4: persistent projection kernel -- GPU tests, A/B of TSDR_PROJ_MODE on one box, bench, launch lists
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -15 gpurun_out/pytest_gpu4.log
AB_ENVS="TSDR_PROJ_MODE=legacy;TSDR_PROJ_MODE=1;TSDR_PROJ_MODE=2" timeout 600 python tools/ab_render.py tempestsdr.jl_b200/libtempest_b200.so > gpurun_out/ab_proj4.log 2>&1
cat gpurun_out/ab_proj4.log
unset TSDR_PROJ_MODE
timeout 900 python bench.py > gpurun_out/bench_default4.json 2> gpurun_out/bench_default4.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default4.err
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches4_cfg3.csv python tools/prof_chain.py cfg3 4 > gpurun_out/launches4_cfg3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches4_cfg5_full.csv python tools/prof_chain.py cfg5 3 full > gpurun_out/launches4_cfg5_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_project|k_beta|k_accumulate' -s 3 -c 3 -o gpurun_out/prof4_aux_cfg3 -f python tools/prof_chain.py cfg3 3 > gpurun_out/prof4_aux_cfg3.log 2>&1
ls -la gpurun_out
