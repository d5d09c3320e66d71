mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e={x['name']:x for x in (d.get('extra') or [])}; rb=e.get('train_denoise_reference_batch',{}); print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'b16', rb.get('b16'), 'b2', rb.get('b2'))"; }
timeout 300 python bench.py --no-cpu-baseline --no-comparator 2>gpurun_out/bench_pdl.err | tee gpurun_out/bench_pdl.json | show "PDL on "
VK_NO_PDL=1 timeout 300 python bench.py --no-cpu-baseline --no-comparator 2>/dev/null | tee gpurun_out/bench_nopdl.json | show "PDL off"
timeout 300 python bench.py --no-cpu-baseline --no-comparator --no-extra 2>/dev/null | show "PDL on again"
