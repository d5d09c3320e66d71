#!/bin/bash
# same-box A/B of an environment switch:  bash tools/gpu_ab.sh VAR=VALUE [rounds]   (A = unset, B = set), alternating
sw="$1"; rounds="${2:-3}"
show() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_ms_per_step']; print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'wgrad', f['conv_wgrad'], 'conv', f['conv_igemm'], 'MHz', d['clocks']['sm_mhz'])"; }
for i in $(seq $rounds); do
  python bench.py --no-cpu-baseline --no-comparator --no-extra --steps 30 2>/dev/null | show "A (default)  "
  env "$sw" python bench.py --no-cpu-baseline --no-comparator --no-extra --steps 30 2>/dev/null | show "B ($sw)"
done
