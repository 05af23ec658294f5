#!/bin/bash
# round-2 GPU call 9 (8 GPUs): all-reduce sweep + the bench at N=8 (C3 weak + C4 strong), guarded by timeouts
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29531 tools/allreduce_sweep.py default > $O/r2c9_ar_default.json 2> $O/r2c9_ar_default.err
NCCL_ALGO=NVLS timeout 240 $T --master-port 29532 tools/allreduce_sweep.py algo_nvls > $O/r2c9_ar_nvls.json 2> $O/r2c9_ar_nvls.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 240 $T --master-port 29533 tools/allreduce_sweep.py info > $O/r2c9_ar_info.json 2> $O/r2c9_ar_info.err
B="bench.py --gpus 8 --steps 50 --warmup 3 --no-ref-ext"
timeout 600 $T --master-port 29534 $B --count-mode sync > $O/r2c9_g8_sync.json 2> $O/r2c9_g8_sync.err
timeout 600 $T --master-port 29535 $B --count-mode async --plain-grad-buffer > $O/r2c9_g8_async_plain.json 2> $O/r2c9_g8_async_plain.err
grep -i "nvls" $O/r2c9_ar_info.err | head -5 > $O/r2c9_nvls_lines.txt
