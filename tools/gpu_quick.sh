#!/bin/bash
# Quick GPU check after a change (about one GPU-minute): per-family breakdown of the denoising and the SISR step and
# the network / kernel / SISR tests.  Output under gpurun_out/quick.log.
mkdir -p gpurun_out
{
  echo "== step_breakdown 32"; timeout 120 python tools/step_breakdown.py 32 | head -16
  echo "== sisr breakdown"; timeout 120 python tools/sisr_train_bench.py 16 bf16 | head -16
  echo "== pytest"; timeout 400 python -m pytest tests/test_gpu_net.py tests/test_gpu_kernels.py tests/test_gpu_sisr.py -x -q 2>&1 | tail -3
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
