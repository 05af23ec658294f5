#!/bin/bash
# round-2 GPU call 15: final verification of the committed code -- whole GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > $O/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c15_pytest.log
tail -3 $O/r2c15_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c15_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c15_smoke.log
tail -2 $O/r2c15_smoke.log
timeout 900 python bench.py > $O/r2c15_bench_default.json 2> $O/r2c15_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c15_bench_reference.json 2> $O/r2c15_bench_reference.err
