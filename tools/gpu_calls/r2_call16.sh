#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29551 bench.py --gpus 4 --steps 50 --warmup 3 --no-ref-ext > $O/r2c16_g4.json 2> $O/r2c16_g4.err
