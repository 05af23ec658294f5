#!/bin/bash
# round-2 GPU call 14 (8 GPUs): C5 tile-occupancy sweep at N=8 (2M Gaussians, 1600^2, one view per GPU + all-reduce)
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
p=29540
for s in 0.5 1 2 4; do
  p=$((p+1))
  timeout 300 $T --master-port $p bench.py --gpus 8 --workload c5 --svec-scale $s --steps 20 --warmup 3 --no-e2e --no-c4-strong --no-ref-ext > $O/r2c14_c5_g8_s$s.json 2> $O/r2c14_c5_g8_s$s.err
done
