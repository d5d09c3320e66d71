#!/bin/bash
# wgrad variants at b=32: round-1 layout vs the 9-tap halo slab with 128 / 112 / 96-pixel K tiles (2 / 3 / 3 stages)
mkdir -p gpurun_out
out=gpurun_out/wgrad_exp.txt
: > $out
for idx in 4 5 6 7; do
  VK_WGRAD_SLAB9=1 timeout 120 python tools/debug_wgrad.py --case $idx >> $out 2>&1 || echo "case $idx FAILED (slab9)" >> $out
done
echo "== round-1 layout" >> $out
timeout 200 python tools/wgrad_bench.py --one 32 0 >> $out 2>&1
for kr in 128 112 96; do
  echo "== slab9 k_rows=$kr" >> $out
  VK_WGRAD_SLAB9=1 timeout 200 python tools/wgrad_bench.py --one 32 $kr >> $out 2>&1
done
cat $out
