#!/bin/bash
# quick check after a change of the parameter-layout kernels: per-family step breakdowns + the network / kernel tests
mkdir -p gpurun_out
{
  for t in 0 1; do
    echo "== VK_PACK_TILED=$t step_breakdown 32"; VK_PACK_TILED=$t timeout 120 python tools/step_breakdown.py 32 | grep -E "total|pack_weights|wgrad_unpack"
  done
  echo "== pytest"; timeout 400 python -m pytest tests/test_gpu_net.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
