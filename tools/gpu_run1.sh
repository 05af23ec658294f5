#!/usr/bin/env bash
# first GPU pass: golden vectors from the reference ext, parity tests, sanitizer smoke, first bench line
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/env.log 2>&1
python - >> gpurun_out/env.log 2>&1 <<'PY'
import torch, os
print(torch.__version__, torch.cuda.get_device_name(0), os.cpu_count())
PY
echo "== golden" ; timeout 600 python tests/golden/make_golden.py > gpurun_out/golden.log 2>&1 ; echo "golden rc=$?"
tail -5 gpurun_out/golden.log
for f in test_ops_gpu test_fused_gpu test_reference_gpu test_golden; do
  echo "== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --tb=short --maxfail=40 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; tail -3 gpurun_out/$f.log
done
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer.log
echo "== bench c1"; timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "rc=$?"; cat gpurun_out/bench_c1.json | head -c 1500
echo "== bench c3"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; cat gpurun_out/bench_c3.json | head -c 3000; tail -5 gpurun_out/bench_c3.err
