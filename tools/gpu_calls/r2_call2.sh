#!/bin/bash
# round-2 GPU call 2: whole GPU suite, bench A/B (SH backward kernels, count modes), ncu captures
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > $O/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c2_pytest.log
tail -3 $O/r2c2_pytest.log
timeout 900 python bench.py > $O/r2c2_bench_default.json 2> $O/r2c2_bench_default.err
B="--no-c4-strong --no-ref-ext --no-cpu-baseline"
GSB200_BWD_SH_VARIANT=1 timeout 600 python bench.py $B > $O/r2c2_bench_bwdv1.json 2> $O/r2c2_bench_bwdv1.err
timeout 600 python bench.py $B --count-mode async > $O/r2c2_bench_async.json 2> $O/r2c2_bench_async.err
timeout 600 python bench.py $B --workload c4 > $O/r2c2_bench_c4.json 2> $O/r2c2_bench_c4.err
# ncu: launch list, then one --set full capture of the four hot kernels of one C3 step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c2_launches_c3.csv python tools/profile_view.py c3 3 > $O/r2c2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_composite|k_preprocess|k_project_bwd_fused" --launch-skip 12 --launch-count 4 -o $O/r2c2_prof_c3 -f python tools/profile_view.py c3 5 > $O/r2c2_ncu_full.log 2>&1
ls -la $O | tail -20
