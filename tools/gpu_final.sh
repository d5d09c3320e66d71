#!/bin/bash
# Final evidence pass of a round: ncu launch list of bench-like steps and a full capture of 24 conv launches.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
   python tools/profile_step.py --steps 3 --batch 32 > gpurun_out/launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_v2 -s 8 -c 24 -f -o gpurun_out/prof_conv \
   python tools/profile_step.py --steps 1 --batch 32 > gpurun_out/prof_conv.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:wgrad_kernel -s 4 -c 6 -f -o gpurun_out/prof_wgrad \
   python tools/profile_step.py --steps 1 --batch 32 > gpurun_out/prof_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
