#!/bin/bash
# sft_bwd: one vs two pixels in flight per thread (VK_SFT_BWD_UNROLL), SISR step breakdown + the SISR tests on both.
mkdir -p gpurun_out
{
  for u in 1 2; do
    echo "== VK_SFT_BWD_UNROLL=$u"
    VK_SFT_BWD_UNROLL=$u timeout 120 python tools/sisr_train_bench.py 16 bf16 | grep -E "workload|sft_bwd|serialised"
    VK_SFT_BWD_UNROLL=$u timeout 120 python tools/sisr_train_bench.py 16 tf32 | grep -E "workload|sft_bwd|serialised"
  done
  echo "== pytest sisr (unroll 2)"; timeout 300 python -m pytest tests/test_gpu_sisr.py tests/test_gpu_modes.py -x -q 2>&1 | tail -3
  echo "== pytest sisr (unroll 1)"; VK_SFT_BWD_UNROLL=1 timeout 300 python -m pytest tests/test_gpu_sisr.py -x -q 2>&1 | tail -3
} > gpurun_out/sft_exp.log 2>&1
cat gpurun_out/sft_exp.log
