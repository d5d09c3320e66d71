#!/bin/bash
# pack_weights: blocks per descriptor (VK_PACK_GRIDX) — per-family step breakdown + the network tests on the candidate
mkdir -p gpurun_out
{
  for g in 296 96 48 24; do
    echo "== VK_PACK_GRIDX=$g"; VK_PACK_GRIDX=$g timeout 60 python tools/step_breakdown.py 32 | grep -E "total|pack_weights"
  done
  echo "== pytest (VK_PACK_GRIDX=48)"; VK_PACK_GRIDX=48 timeout 100 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -2
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
