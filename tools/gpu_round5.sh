#!/bin/bash
# round 2, call 5: k_project_p with tiled TMA -- targeted GPU tests, A/B of TSDR_PROJ_MODE, launch lists, ncu
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q -k "chain or vsync or fullres or golden or cfg1 or search or block_integration" > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
tail -15 gpurun_out/pytest_gpu5.log
AB_ENVS="TSDR_PROJ_MODE=legacy;TSDR_PROJ_MODE=1;TSDR_PROJ_MODE=2;TSDR_PROJ_MODE=3" timeout 600 python tools/ab_render.py tempestsdr.jl_b200/libtempest_b200.so > gpurun_out/ab_proj5.log 2>&1
cat gpurun_out/ab_proj5.log
unset TSDR_PROJ_MODE
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/launches5_cfg3.csv python tools/prof_chain.py cfg3 4 > gpurun_out/launches5_cfg3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/launches5_cfg5_full.csv python tools/prof_chain.py cfg5 3 full > gpurun_out/launches5_cfg5_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_project' -s 1 -c 1 -o gpurun_out/prof5_project_cfg3 -f python tools/prof_chain.py cfg3 3 > gpurun_out/prof5_project_cfg3.log 2>&1
ls -la gpurun_out
