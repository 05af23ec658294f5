#!/bin/bash
# round-2 final GPU call (1 GPU, ~6 min of box time left): the whole GPU suite of the committed code incl. the new
# gsb200_knn and drop-in tests (no -x: every failure is wanted), smoke(), the default bench line.
mkdir -p gpurun_out
O=gpurun_out
export GSB200_TEST_NOTES=$PWD/$O/r2final_test_notes.txt
rm -f $GSB200_TEST_NOTES
timeout 240 python -m pytest tests -q -m gpu --timeout 200 --durations=12 -p no:cacheprovider > $O/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2final_pytest.log
grep -v "^out_rgb\|^\[DEBUG\]" $O/r2final_pytest.log | tail -40
cat $GSB200_TEST_NOTES 2>/dev/null
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2final_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2final_smoke.log
grep -v "^out_rgb" $O/r2final_smoke.log | tail -2
timeout 300 python bench.py > $O/r2final_bench_default.json 2> $O/r2final_bench_default.err; echo "bench rc=$?"
wc -l $O/r2final_bench_default.json
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2final_bench_default.json') if l.startswith('{')][-1])
    print("ms/step %.4f runs %s e2e %.4f" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]], d["e2e"]["ms_per_step"]))
    print({k:round(v["ms"],4) for k,v in d["stages"].items()}, d["clocks"])
except Exception as e: print("bench parse failed", e)
PY
tail -3 $O/r2final_bench_default.err
