#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== golden" ; timeout 900 python tests/golden/make_golden.py > gpurun_out/golden.log 2>&1 ; echo "golden rc=$?"; grep "^golden" gpurun_out/golden.log; ls -la gpurun_out/golden/ 2>/dev/null
for f in test_ops_gpu test_fused_gpu test_reference_gpu test_golden test_properties_gpu; do
  echo "== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; tail -3 gpurun_out/$f.log
done
echo "== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_c3.json'))
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "value", d["value"])
    print({k:round(v["ms"],4) for k,v in d["stages"].items()}, d["view_stats"], d["clocks"])
    print(d["cpu_baseline"])
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_c3.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_composite -s 8 -c 2 -o gpurun_out/prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
