#!/bin/bash
# round-2 GPU call 6: whole GPU suite + full default bench (e2e, c4_strong, reference ext)
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > $O/r2c6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c6_pytest.log
tail -4 $O/r2c6_pytest.log
timeout 900 python bench.py > $O/r2c6_bench_default.json 2> $O/r2c6_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c6_bench_reference.json 2> $O/r2c6_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c6_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c6_smoke.log
