#!/bin/bash
# round-2 GPU call 5 (2 GPUs): replica-consistency test, bench under torchrun at N=2 (C3 weak + C4 strong sub-record)
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_parity_fullsize_gpu.py -q -m gpu --timeout 300 -k "two_gpu or c5" > $O/r2c5_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c5_pytest.log
tail -4 $O/r2c5_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > $O/r2c5_bench_g2.json 2> $O/r2c5_bench_g2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --plain-grad-buffer --no-e2e > $O/r2c5_bench_g2_plain.json 2> $O/r2c5_bench_g2_plain.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r2c5_bench_ref_g2.json 2> $O/r2c5_bench_ref_g2.err
tail -2 $O/r2c5_bench_g2.err
