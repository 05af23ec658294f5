#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > $O/r2c12_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c12_pytest.log
tail -3 $O/r2c12_pytest.log
B="--no-c4-strong --no-ref-ext --no-cpu-baseline --no-e2e --steps 30"
for wl in c3 c4 c5; do
  timeout 600 python bench.py $B --workload $wl > $O/r2c12_bench_cull_$wl.json 2> $O/r2c12_bench_cull_$wl.err
done
