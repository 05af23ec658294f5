#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_composite -s 8 -c 2 -o gpurun_out/prof_c3_r4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
