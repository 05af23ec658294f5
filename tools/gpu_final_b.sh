#!/usr/bin/env bash
# final 2-GPU validation of the torchrun path (rank-agreed warm-up, NCCL all-reduce)
set -u
mkdir -p gpurun_out
N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_c3_g$N.json 2> gpurun_out/bench_c3_g$N.err; echo "c3 rc=$?"
python - gpurun_out/bench_c3_g$N.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print("n_gpus",d["n_gpus"],"ms/step %.4f runs %s value %.4g e2e %.4f %s" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_all_runs"]))
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_g$N.json 2> gpurun_out/bench_c4_g$N.err; echo "c4 rc=$?"
python - gpurun_out/bench_c4_g$N.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print("c4 n_gpus",d["n_gpus"],"ms/step %.4f runs %s value %.4g views/gpu %s" % (d["ms_per_step"], [round(x,3) for x in d["ms_per_step_all_runs"]], d["value"], d["config"]["views_per_gpu"]))
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --impl reference --steps 1 --warmup 0 --workload c1 | head -c 300; echo
tail -2 gpurun_out/bench_c3_g$N.err
