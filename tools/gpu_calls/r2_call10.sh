#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
GSB200_FWD_SH_VARIANT=1 timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_reference_gpu.py tests/test_parity_fullsize_gpu.py tests/test_properties_gpu.py -q -m gpu --timeout 900 > $O/r2c10_pytest_fwd2.log 2>&1; echo "pytest rc=$?" >> $O/r2c10_pytest_fwd2.log
tail -3 $O/r2c10_pytest_fwd2.log
B="--no-c4-strong --no-ref-ext --no-cpu-baseline --no-e2e --steps 30"
GSB200_FWD_SH_VARIANT=1 timeout 600 python bench.py $B > $O/r2c10_bench_fwd2.json 2> $O/r2c10_bench_fwd2.err
for v in fwd2_mb5 fwd2_b32 fwd2_mb4; do
  GSB200_FWD_SH_VARIANT=1 GSB200_LIB=$PWD/gsgen_b200/_variants/lib_$v.so timeout 600 python bench.py $B > $O/r2c10_bench_$v.json 2> $O/r2c10_bench_$v.err
done
GSB200_FWD_SH_VARIANT=1 timeout 600 python bench.py $B --workload c4 > $O/r2c10_bench_fwd2_c4.json 2> $O/r2c10_bench_fwd2_c4.err
GSB200_FWD_SH_VARIANT=1 timeout 600 python bench.py $B --workload c5 > $O/r2c10_bench_fwd2_c5.json 2> $O/r2c10_bench_fwd2_c5.err
timeout 600 python bench.py $B --workload c5 > $O/r2c10_bench_fwd1_c5.json 2> $O/r2c10_bench_fwd1_c5.err
GSB200_FWD_SH_VARIANT=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_composite_fwd_sh2" --launch-skip 3 --launch-count 1 -o $O/r2c10_prof_fwd2 -f python tools/profile_view.py c3 5 > $O/r2c10_ncu.log 2>&1
