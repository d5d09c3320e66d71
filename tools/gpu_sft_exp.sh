#!/bin/bash
# SISR step breakdown after the small-kernel changes (sft_bwd, ca_layer, knet_head) + the SISR tests.
mkdir -p gpurun_out
{
  echo "== bench"; timeout 120 python tools/sisr_train_bench.py 16 bf16 | grep -E "workload|sft_bwd|serialised|ca_layer|knet_head|elbo"
  echo "== pytest sisr"; timeout 300 python -m pytest tests/test_gpu_sisr.py tests/test_gpu_modes.py tests/test_gpu_sisr_loss.py -x -q 2>&1 | tail -3
} > gpurun_out/sft_exp2.log 2>&1
cat gpurun_out/sft_exp2.log
