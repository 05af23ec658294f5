#!/bin/bash
# round-2 GPU call 17 (2 GPUs): sparse all-reduce -- 2-GPU tests, then the bench sparse vs dense under torchrun
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_fused_gpu.py tests/test_train_step_gpu.py -q -m gpu --timeout 300 > $O/r2c17_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c17_pytest.log
tail -3 $O/r2c17_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 50 --warmup 3 --no-ref-ext"
timeout 900 $T --master-port 29561 $B > $O/r2c17_g2_sparse.json 2> $O/r2c17_g2_sparse.err
timeout 900 $T --master-port 29562 $B --dense-allreduce --no-e2e > $O/r2c17_g2_dense.json 2> $O/r2c17_g2_dense.err
