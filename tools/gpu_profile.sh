#!/usr/bin/env bash
# ncu evidence for profiles/ (one GPU; never a multi-rank command): the launch list of the bench command and one
# `--set full` capture of the two SH composite kernels.   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r2'
set -u
TAG=${1:-rN}
mkdir -p gpurun_out
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu --set full of the composite kernels (launches 8, 9 = one forward + one backward after warm-up)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_composite -s 8 -c 2 \
  -o gpurun_out/${TAG}_ncu_full_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | grep "${TAG}_"
