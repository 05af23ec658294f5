#!/bin/bash
# round-2 GPU call 13: suite at north_star tolerances, smoke, ncu captures of the final kernels, default bench
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > $O/r2c13_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c13_pytest.log
tail -3 $O/r2c13_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c13_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c13_smoke.log
tail -2 $O/r2c13_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c13_launches_c3.csv python tools/profile_view.py c3 3 > $O/r2c13_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_composite|k_preprocess|k_project_bwd_fused|k_emit_tiles|k_tile_ranges" --launch-skip 18 --launch-count 6 -o $O/r2c13_prof_c3 -f python tools/profile_view.py c3 5 > $O/r2c13_ncu_full.log 2>&1
timeout 900 python bench.py > $O/r2c13_bench_default.json 2> $O/r2c13_bench_default.err
