#!/bin/bash
# round-2 GPU call 8 (2 GPUs): polled async-count mode vs synchronous under torchrun, registered vs plain buffer
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_round2_gpu.py -q -m gpu --timeout 300 > $O/r2c8_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c8_pytest.log
tail -3 $O/r2c8_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 50 --warmup 3 --no-ref-ext"
timeout 900 $T --master-port 29521 $B --count-mode sync > $O/r2c8_g2_sync.json 2> $O/r2c8_g2_sync.err
timeout 900 $T --master-port 29522 $B --count-mode async > $O/r2c8_g2_async.json 2> $O/r2c8_g2_async.err
timeout 900 $T --master-port 29523 $B --count-mode async --plain-grad-buffer > $O/r2c8_g2_async_plain.json 2> $O/r2c8_g2_async_plain.err
timeout 900 $T --master-port 29524 $B --count-mode sync --plain-grad-buffer > $O/r2c8_g2_sync_plain.json 2> $O/r2c8_g2_sync_plain.err
