#!/bin/bash
# compute-sanitizer over the full-resolution chain's kernels (k_render_full, k_project_p<true>, k_beta<true>, k_accumulate_rows)
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 60 python -m pytest tests -m gpu -x -q -k "fullres" > gpurun_out/pytest_fullres_last.log 2>&1; tail -2 gpurun_out/pytest_fullres_last.log
timeout 100 $CS --tool racecheck python tools/sanitize.py fullres > gpurun_out/racecheck_fullres.txt 2>&1
grep -E "RACECHECK SUMMARY|sanitize.py done" gpurun_out/racecheck_fullres.txt
