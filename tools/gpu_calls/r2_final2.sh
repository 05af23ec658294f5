#!/bin/bash
# round-2 final GPU call #2 (1 GPU): whole GPU suite of the committed code (drop-in test with the tie-order handling),
# smoke(), the default bench line (count mode auto -> async at N=1) and a short sync-mode run beside it.
mkdir -p gpurun_out
O=gpurun_out
export GSB200_TEST_NOTES=$PWD/$O/r2final2_test_notes.txt
rm -f $GSB200_TEST_NOTES
timeout 200 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider > $O/r2final2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2final2_pytest.log
grep -v "^out_rgb\|^\[DEBUG\]" $O/r2final2_pytest.log | tail -8
cat $GSB200_TEST_NOTES 2>/dev/null
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2final2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2final2_smoke.log
grep -v "^out_rgb" $O/r2final2_smoke.log | tail -2
timeout 300 python bench.py > $O/r2final2_bench_default.json 2> $O/r2final2_bench_default.err; echo "bench rc=$? stdout lines: $(wc -l < $O/r2final2_bench_default.json)"
timeout 120 python bench.py --count-mode sync --no-cpu-baseline --no-ref-ext --no-c4-strong --no-e2e > $O/r2final2_bench_sync.json 2> $O/r2final2_bench_sync.err; echo "bench sync rc=$?"
python - <<'PY'
import json
for f in ("r2final2_bench_default", "r2final2_bench_sync"):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, d["config"]["count_mode"], "ms/step %.4f runs %s" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]]), "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step_all_runs"))
        print("  ", {k:round(v["ms"],4) for k,v in d["stages"].items()}, "overflows", d.get("tile_list_overflows"))
    except Exception as e: print(f, "parse failed", e)
PY
