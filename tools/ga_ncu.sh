#!/bin/bash
# per-kernel durations of GA generations at the cfg4 shape: plain run first, then the ncu launch list of the same command
set -e
mkdir -p gpurun_out
GA_GENS=8 python tools/time_ga.py > gpurun_out/ga_plain.txt 2>&1
GA_GENS=8 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/ga_launches.csv python tools/time_ga.py > gpurun_out/ga_ncu.log 2>&1 || true
