#!/bin/bash
# Round-2 check of the deterministic super-resolution step: bitwise repeatability of every engine buffer, the SISR
# tests, and what the mode costs.  Output under gpurun_out/.
mkdir -p gpurun_out
{
  echo "== det_check_sr 4 bf16 4"; timeout 200 python tools/det_check_sr.py 4 bf16 4
  echo "== pytest sisr"; timeout 400 python -m pytest tests/test_gpu_sisr.py tests/test_gpu_sisr_loss.py -x -q 2>&1 | tail -15
  echo "== bench atomic"; timeout 120 python tools/sisr_train_bench.py 16 bf16 > gpurun_out/sisr_atomic.txt; head -2 gpurun_out/sisr_atomic.txt
  echo "== bench det"; timeout 120 python tools/sisr_train_bench.py 16 bf16 --det > gpurun_out/sisr_det.txt; head -14 gpurun_out/sisr_det.txt
} > gpurun_out/det_sr.log 2>&1
tail -40 gpurun_out/det_sr.log
