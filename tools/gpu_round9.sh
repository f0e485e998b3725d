#!/bin/bash
# round 2, call 13: tightened k_accumulate_rows + programmatic dependent launch in the three-level FFT
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests -m gpu -x -q -k "fullres or autocorr or extract_configuration or findmax_device or Spectrum or Welch or resampler" > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu9.log
tail -4 gpurun_out/pytest_gpu9.log
for rep in 1 2; do for k in 23 24 26; do
TSDR_FFT_PDL=0 python tools/run_autocorr.py $k 50 2>&1 | sed 's/^/pdl=0 /'
TSDR_FFT_PDL=1 python tools/run_autocorr.py $k 50 2>&1 | sed 's/^/pdl=1 /'
done; done > gpurun_out/ab_fft_pdl9.log 2>&1
cat gpurun_out/ab_fft_pdl9.log
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 16 --csv --log-file gpurun_out/launches9_cfg5_full.csv python tools/prof_chain.py cfg5 2 full > gpurun_out/launches9_cfg5_full.log 2>&1
timeout 200 python bench.py --workload cfg5_fullres --steps 2 --warmup 3 > gpurun_out/bench_cfg5_fullres_n1.json 2> gpurun_out/bench_cfg5_fullres_n1.err; echo "rc=$?"
