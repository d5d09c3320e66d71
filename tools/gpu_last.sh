#!/bin/bash
# Last GPU call of a round on a short budget: the whole -m gpu suite, smoke(), one default bench line.
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_last.json
