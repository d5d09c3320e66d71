#!/bin/bash
# 2-GPU experiment: what the gradient exchange costs (same box: N=1, then N=2 with the overlapped buckets, the single
# blocking all-reduce, and NCCL restricted to a few CTAs)
mkdir -p gpurun_out
out=gpurun_out/n2_exp.txt
: > $out
B="bench.py --no-cpu-baseline --no-comparator --no-extra --steps 30 --warmup 5"
run1() { python $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['families_ms_per_step'].get('wgrad_unpack'))" >> $out; }
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 $B --gpus 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['families_ms_per_step'].get('wgrad_unpack'))" >> $out; }
run1 "N=1"
run2 "N=2 overlap" 29601
VIRNET_B200_OVERLAP_ALLREDUCE=0 run2 "N=2 blocking" 29602
NCCL_MAX_CTAS=4 run2 "N=2 overlap NCCL_MAX_CTAS=4" 29603
NCCL_MAX_CTAS=4 VIRNET_B200_OVERLAP_ALLREDUCE=0 run2 "N=2 blocking NCCL_MAX_CTAS=4" 29604
NCCL_MAX_CTAS=16 run2 "N=2 overlap NCCL_MAX_CTAS=16" 29605
run1 "N=1 again"
cat $out
