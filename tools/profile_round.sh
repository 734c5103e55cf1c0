#!/bin/bash
# Evidence of a round (run on the GPU box through gpurun; everything lands in gpurun_out/):
#   1. plain runs first (a number printed under ncu is never a bench value);
#   2. launch list of a short bench.py run (gpu__time_duration per launch: cold, serialised -- compare shares);
#   3. one `ncu --set full` capture of tools/prof_target.py (one launch of every kernel of the path).
set -e
R=${1:-r02}
mkdir -p gpurun_out
python tools/prof_target.py > gpurun_out/${R}_prof_target_plain.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_short.json 2> gpurun_out/${R}_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_launches_ncu.log 2>&1 || true
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${R}_prof_target \
    python tools/prof_target.py > gpurun_out/${R}_prof_target_ncu.log 2>&1 || true
tail -2 gpurun_out/${R}_prof_target_ncu.log
# reduce on the box (gpurun brings back at most 64 MiB): the summary always, the report itself only when it is small enough
python tools/ncu_summary.py gpurun_out/${R}_prof_target.ncu-rep gpurun_out/${R}_ncu_summary.json > /dev/null
ncu -i gpurun_out/${R}_prof_target.ncu-rep --page raw --csv | gzip -9 > gpurun_out/${R}_prof_target_raw.csv.gz
if [ $(stat -c %s gpurun_out/${R}_prof_target.ncu-rep) -gt 40000000 ]; then rm gpurun_out/${R}_prof_target.ncu-rep; fi
du -sh gpurun_out
