#!/bin/bash
# ncu --set full with source attribution of ONE 96-channel conv launch per epilogue flavour (b=32, 128x128)
mkdir -p gpurun_out
for epi in resid_dual out2; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 1 -c 1 -f -o gpurun_out/prof_c96_$epi \
   python tools/one_conv.py --c 96 --h 128 --n 32 --epi $epi > gpurun_out/prof_c96_$epi.log 2>&1; echo "ncu $epi rc=$?"
done
timeout 300 python tools/conv_bench.py --n 32 > gpurun_out/conv_bench.txt 2>&1; tail -30 gpurun_out/conv_bench.txt
