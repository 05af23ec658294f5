#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== default threads c3"; python tools/trace_step.py c3 0 2>&1 | tail -12
echo "== 1 thread c3"; python tools/trace_step.py c3 1 2>&1 | tail -10
echo "== 1 thread c2"; python tools/trace_step.py c2 1 2>&1 | tail -10
