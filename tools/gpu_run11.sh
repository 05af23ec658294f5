#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for f in test_ops_gpu test_fused_gpu test_reference_gpu test_golden test_properties_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; tail -1 gpurun_out/$f.log
done
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.4f runs %s (with stage events %.4f) e2e %.4f value %.4g" % (d["ms_per_step"], [round(x,3) for x in d.get("ms_per_step_all_runs",[])], d.get("ms_per_step_with_stage_events",-1), d["e2e"]["ms_per_step"] if d.get("e2e") else -1, d["value"]))
    print({k:round(v["ms"],4) for k,v in d["stages"].items()}, "roofline frac %.4f" % d["roofline"]["frac"], d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
except Exception as e: print("bench parse failed", e)
PY
}
echo "== bench c3"; timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; summ gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
echo "== vs reference ext"; timeout 900 python tools/compare_reference_ext.py > gpurun_out/vs_reference_ext.json 2> gpurun_out/vs_reference_ext.err; echo "rc=$?"; tail -3 gpurun_out/vs_reference_ext.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/vs_reference_ext.json'))
    for c in ("c2","c3"):
        print(c, d[c]["config"], "D", d[c]["N_with_dub"], "fwd diff", d[c]["fwd_max_abs_diff"])
        for k,v in d[c].items():
            if isinstance(v,dict): print("   ",k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print("parse failed", e)
PY
