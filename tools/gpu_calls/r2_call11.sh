#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
GSB200_FWD_SH_VARIANT=1 timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_reference_gpu.py tests/test_parity_fullsize_gpu.py tests/test_properties_gpu.py -q -m gpu --timeout 900 > $O/r2c11_pytest_fwdh.log 2>&1; echo "pytest rc=$?" >> $O/r2c11_pytest_fwdh.log
tail -3 $O/r2c11_pytest_fwdh.log
B="--no-c4-strong --no-ref-ext --no-cpu-baseline --no-e2e --steps 30"
for wl in c3 c4 c5; do
  GSB200_FWD_SH_VARIANT=1 timeout 600 python bench.py $B --workload $wl > $O/r2c11_bench_fwdh_$wl.json 2> $O/r2c11_bench_fwdh_$wl.err
done
GSB200_FWD_SH_VARIANT=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_composite_fwd_shh" --launch-skip 3 --launch-count 1 -o $O/r2c11_prof_fwdh -f python tools/profile_view.py c3 5 > $O/r2c11_ncu.log 2>&1
