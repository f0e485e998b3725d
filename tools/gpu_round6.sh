#!/bin/bash
# round 2, call 6: k_project_p (tiled TMA, fold in its own kernel) -- targeted GPU tests, A/B of modes and ring depths, ncu
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q -k "chain or vsync or fullres or golden or cfg1 or search or block_integration" > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
tail -5 gpurun_out/pytest_gpu6.log
L=tempestsdr.jl_b200
AB_ENVS="TSDR_PROJ_MODE=legacy;TSDR_PROJ_MODE=1;TSDR_PROJ_MODE=2;TSDR_PROJ_MODE=3;TSDR_PROJ_MODE=4" timeout 900 python tools/ab_render.py $L/libtempest_b200.so $L/libtempest_b200_s2.so $L/libtempest_b200_s4.so > gpurun_out/ab_proj6.log 2>&1
cat gpurun_out/ab_proj6.log
unset TSDR_PROJ_MODE
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/launches6_cfg3.csv python tools/prof_chain.py cfg3 4 > gpurun_out/launches6_cfg3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/launches6_cfg2.csv python tools/prof_chain.py cfg2 4 > gpurun_out/launches6_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_project|k_fold' -s 2 -c 2 -o gpurun_out/prof6_project_cfg3 -f python tools/prof_chain.py cfg3 3 > gpurun_out/prof6_project_cfg3.log 2>&1
ls -la gpurun_out
