#!/bin/bash
# One GPU call that re-establishes the measured state of the repo: parity tests, smoke, both bench arms, the SISR /
# inference / graph side benches, and the ncu evidence pass (tools/gpu_final.sh).  ~8 GPU-minutes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/sisr_train_bench.py 16 bf16 > gpurun_out/sisr_step.txt 2>&1; head -1 gpurun_out/sisr_step.txt
timeout 300 python tools/infer_bench.py > gpurun_out/infer.jsonl 2>&1; tail -4 gpurun_out/infer.jsonl
timeout 300 python tools/graph_bench.py 2 4 16 > gpurun_out/graph.jsonl 2>&1; tail -3 gpurun_out/graph.jsonl
timeout 300 python tools/aux_kernels_bench.py > gpurun_out/aux.jsonl 2>&1; tail -6 gpurun_out/aux.jsonl
bash tools/gpu_final.sh
