#!/usr/bin/env bash
# multi-GPU bench: $1 = number of GPUs.  Inner timeouts are short on purpose: a multi-rank hang is charged N x box time.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_c3_g$N.json 2> gpurun_out/bench_c3_g$N.err; echo "rc=$?"
tail -3 gpurun_out/bench_c3_g$N.err
python - gpurun_out/bench_c3_g$N.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print("n_gpus",d["n_gpus"],"ms/step %.4f runs %s value %.4g e2e %.4f" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]], d["value"], d["e2e"]["ms_per_step"]), d["config"]["grad_allreduce_bytes"])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_g$N.json 2> gpurun_out/bench_c4_g$N.err; echo "rc=$?"
python - gpurun_out/bench_c4_g$N.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print("c4 n_gpus",d["n_gpus"],"ms/step %.4f runs %s value %.4g" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]], d["value"]), d["config"]["views_per_gpu"])
PY
