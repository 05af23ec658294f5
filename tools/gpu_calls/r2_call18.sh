#!/bin/bash
# round-2 GPU call 18 (4 GPUs): sparse all-reduce v2 -- 2-GPU tests, bench at N=2 and N=4 (sparse), N=4 dense for comparison
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_round2_gpu.py -q -m gpu --timeout 300 -k "two_gpu" > $O/r2c18_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c18_pytest.log
tail -3 $O/r2c18_pytest.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --steps 50 --warmup 3 --no-ref-ext --no-e2e"
timeout 600 $T --nproc-per-node 2 --master-port 29571 $B --gpus 2 > $O/r2c18_g2_sparse.json 2> $O/r2c18_g2_sparse.err
timeout 600 $T --nproc-per-node 4 --master-port 29572 $B --gpus 4 > $O/r2c18_g4_sparse.json 2> $O/r2c18_g4_sparse.err
timeout 600 $T --nproc-per-node 4 --master-port 29573 $B --gpus 4 --dense-allreduce --no-c4-strong > $O/r2c18_g4_dense.json 2> $O/r2c18_g4_dense.err
