#!/usr/bin/env bash
# single-GPU validation: what the driver runs at round end (tests, smoke, both bench arms)   gpurun --timeout 1500 -- bash tools/gpu_validate.sh
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench (defaults)"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print("ms/step %.4f runs %s e2e %.4f value %.4g roofline %.4f traffic %s" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]], d["e2e"]["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["traffic"]))
    print({k:round(v["ms"],4) for k,v in d["stages"].items()}, d["clocks"])
    print(d["cpu_baseline"])
except Exception as e: print("bench parse failed", e)
PY
echo "== bench --impl reference (defaults)"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_default.json 2> gpurun_out/bench_reference_default.err; echo "rc=$?"; head -c 700 gpurun_out/bench_reference_default.json
