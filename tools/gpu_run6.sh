#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.4f (with stage events %.4f) e2e %.4f  clocks %s" % (d["ms_per_step"], d.get("ms_per_step_with_stage_events",-1), d["e2e"]["ms_per_step"] if d.get("e2e") else -1, d["clocks"]))
except Exception as e: print("bench parse failed", e)
PY
}
for per in 0 0.01 0.1 0.25 0; do
  echo "== c3 clock-period $per"; timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --clock-period $per > gpurun_out/b_$per.json 2> gpurun_out/b_$per.err; summ gpurun_out/b_$per.json
done
echo "== c2 clock-period 0";  timeout 600 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline --clock-period 0 > gpurun_out/b2_0.json 2>/dev/null; summ gpurun_out/b2_0.json
echo "== c2 clock-period 0.1";  timeout 600 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu-baseline --clock-period 0.1 > gpurun_out/b2_1.json 2>/dev/null; summ gpurun_out/b2_1.json
