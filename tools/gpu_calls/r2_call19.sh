#!/bin/bash
# round-2 GPU call 19 (2 GPUs): sparse all-reduce v2 -- 2-GPU tests, bench at N=2 sparse vs dense
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_round2_gpu.py -q -m gpu --timeout 300 -k "two_gpu" > $O/r2c19_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c19_pytest.log
tail -3 $O/r2c19_pytest.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
B="bench.py --steps 50 --warmup 3 --no-ref-ext --no-e2e --gpus 2"
timeout 600 $T --master-port 29581 $B > $O/r2c19_g2_sparse.json 2> $O/r2c19_g2_sparse.err
timeout 600 $T --master-port 29582 $B --dense-allreduce --no-c4-strong > $O/r2c19_g2_dense.json 2> $O/r2c19_g2_dense.err
