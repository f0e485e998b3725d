#!/bin/bash
# one gpurun call: GPU tests, the default bench, ncu launch lists and full captures (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3.csv python tools/prof_chain.py cfg3 4 > gpurun_out/launches_cfg3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg2.csv python tools/prof_chain.py cfg2 4 > gpurun_out/launches_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/prof_render_cfg3 -f python tools/prof_chain.py cfg3 4 > gpurun_out/prof_render_cfg3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/prof_render_cfg2 -f python tools/prof_chain.py cfg2 4 > gpurun_out/prof_render_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/prof_render_cfg3_i16 -f python tools/prof_chain.py cfg3 4 i16 > gpurun_out/prof_render_cfg3_i16.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_project|k_beta|k_accumulate' -s 3 -c 3 -o gpurun_out/prof_aux_cfg3 -f python tools/prof_chain.py cfg3 3 > gpurun_out/prof_aux_cfg3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_ -s 10 -c 5 -o gpurun_out/prof_fft3 -f python tools/run_autocorr.py 24 4 > gpurun_out/prof_fft3.log 2>&1
ls -la gpurun_out
