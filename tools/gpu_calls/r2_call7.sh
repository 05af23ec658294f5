#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B="--no-c4-strong --no-ref-ext --no-cpu-baseline --no-e2e --steps 30"
for v in bwd_nw4 fwd_nw4 fwd_nw4_b32 both_nw4; do
  GSB200_LIB=$PWD/gsgen_b200/_variants/lib_$v.so timeout 600 python bench.py $B > $O/r2c7_bench_$v.json 2> $O/r2c7_bench_$v.err
done
GSB200_LIB=$PWD/gsgen_b200/_variants/lib_both_nw4.so timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_round2_gpu.py -q -m gpu -k "sh or kernels_agree" > $O/r2c7_pytest_nw4.log 2>&1
tail -2 $O/r2c7_pytest_nw4.log
