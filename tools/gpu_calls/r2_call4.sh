#!/bin/bash
# round-2 GPU call 4: flush v4 A/B + failing tests re-run + ncu of the backward
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_round2_gpu.py tests/test_parity_fullsize_gpu.py tests/test_reference_gpu.py tests/test_fused_gpu.py -q -m gpu --timeout 900 > $O/r2c4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c4_pytest.log
tail -5 $O/r2c4_pytest.log
B="--no-c4-strong --no-ref-ext --no-cpu-baseline --no-e2e --steps 30"
timeout 600 python bench.py $B > $O/r2c4_bench_v4.json 2> $O/r2c4_bench_v4.err
for v in bwd_nopf bwd_mb2; do
  GSB200_LIB=$PWD/gsgen_b200/_variants/lib_$v.so timeout 600 python bench.py $B > $O/r2c4_bench_$v.json 2> $O/r2c4_bench_$v.err
done
timeout 600 python bench.py $B --workload c4 > $O/r2c4_bench_c4.json 2> $O/r2c4_bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_composite_bwd_sh" --launch-skip 3 --launch-count 1 -o $O/r2c4_prof_bwdsh -f python tools/profile_view.py c3 5 > $O/r2c4_ncu_full.log 2>&1
