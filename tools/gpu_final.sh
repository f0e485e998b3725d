#!/bin/bash
# final call of the round: full GPU test suite, default bench + reference arm, launch lists and full ncu captures of this round's kernels
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -4 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_default_final.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/launches_final_cfg3.csv python tools/prof_chain.py cfg3 4 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_accumulate_rows|k_project_p|k_beta' -s 3 -c 3 -o gpurun_out/prof_final_fullres_aux -f python tools/prof_chain.py cfg5 2 full > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_ -s 10 -c 5 -o gpurun_out/prof_final_fft3 -f python tools/run_autocorr.py 24 4 > /dev/null 2>&1
ls -la gpurun_out | tail -12
