#!/bin/bash
# round-2 GPU call 1: full-size parity tests + C5 occupancy sweep + C4 single GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_env.log 2>&1
timeout 1500 python -m pytest tests/test_parity_fullsize_gpu.py tests/test_reference_gpu.py -q -m gpu --timeout 900 > gpurun_out/r2_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_parity.log
for s in 0.5 1 2 4; do
  timeout 600 python bench.py --workload c5 --svec-scale $s --no-cpu-baseline --steps 20 > gpurun_out/r2_c5_s$s.json 2> gpurun_out/r2_c5_s$s.err
done
timeout 600 python bench.py --workload c4 --no-cpu-baseline --steps 20 > gpurun_out/r2_c4_g1.json 2> gpurun_out/r2_c4_g1.err
tail -5 gpurun_out/r2_parity.log
