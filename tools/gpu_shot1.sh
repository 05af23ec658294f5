#!/usr/bin/env bash
# minimal GPU run: only the tests added since the last full GPU run + the aux measurement
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest "tests/test_fused_gpu.py::test_raw_params_equal_torch_activations" tests/test_optim_gpu.py tests/test_train_step_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/shot1_pytest.log
cat gpurun_out/shot1_pytest.log | tail -30
timeout 70 python tools/bench_aux.py > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err; echo "aux rc=$?"
cat gpurun_out/bench_aux.json | head -40; tail -3 gpurun_out/bench_aux.err
