#!/bin/bash
# first GPU bring-up: each stage in its own process under a hard timeout so that a hung kernel cannot eat the lease
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for prec in simt tf32 fp32; do
  echo "=== smoke $prec" | tee -a gpurun_out/smoke.log
  B2J_SMOKE_PRECISIONS=$prec timeout -s KILL 240 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.log 2>&1
  echo "exit $?" | tee -a gpurun_out/smoke.log
done
tail -30 gpurun_out/smoke.log
timeout -s KILL 900 python -m pytest tests/test_function.py tests/test_random.py tests/test_reduce_window.py -m gpu -q --timeout 120 -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_small.log
timeout -s KILL 1500 python -m pytest tests/test_basic_ops.py -m gpu -q --timeout 120 2>&1 | tail -60 | tee gpurun_out/pytest_basic.log
timeout -s KILL 1500 python -m pytest tests/test_conv.py -m gpu -q --timeout 120 2>&1 | tail -60 | tee gpurun_out/pytest_conv.log
