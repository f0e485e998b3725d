#!/bin/bash
# round 2, call 7: A/B of k_render CTA sizes (160 / 200 / 320 / 400 threads) on one box
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
L=tempestsdr.jl_b200
timeout 900 python tools/ab_render.py $L/libtempest_b200.so $L/libtempest_b200_t128.so $L/libtempest_b200_t96.so $L/libtempest_b200_t80.so > gpurun_out/ab_threads7.log 2>&1
cat gpurun_out/ab_threads7.log
