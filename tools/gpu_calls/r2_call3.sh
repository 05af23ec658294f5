#!/bin/bash
# round-2 GPU call 3: whole GPU suite (no -x), SH backward v3 A/B, forward/backward tuning variants, ncu of the backward
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > $O/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c3_pytest.log
tail -5 $O/r2c3_pytest.log
B="--no-c4-strong --no-ref-ext --no-cpu-baseline --no-e2e --steps 30"
timeout 600 python bench.py $B > $O/r2c3_bench_v3.json 2> $O/r2c3_bench_v3.err
GSB200_BWD_SH_VARIANT=1 timeout 600 python bench.py $B > $O/r2c3_bench_bwdv1.json 2> $O/r2c3_bench_bwdv1.err
for v in fwd_mb6 fwd_mb1 fwd_mb4 fwd_b32 fwd_b32_mb6 bwd_mb2; do
  GSB200_LIB=$PWD/gsgen_b200/_variants/lib_$v.so timeout 600 python bench.py $B > $O/r2c3_bench_$v.json 2> $O/r2c3_bench_$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_composite_bwd_sh" --launch-skip 3 --launch-count 1 -o $O/r2c3_prof_bwdsh -f python tools/profile_view.py c3 5 > $O/r2c3_ncu_full.log 2>&1
ls -la $O | tail -5
