#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for bench.py; run under gpurun, outputs land in gpurun_out/.
# usage: profiles/run_ncu.sh <tag> [rows]
TAG=${1:-r01}
ROWS=${2:-800000000}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity --rows $ROWS"
# 1. every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
# 2. full capture of the two dominant kernels (second launch of each = warm)
ncu --set full --clock-control none --import-source on -k regex:scatter_rows_kernel -s 4 -c 2 \
    -f -o gpurun_out/${TAG}_scatter $CMD > gpurun_out/${TAG}_scatter.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bucket_join_kernel -s 1 -c 1 \
    -f -o gpurun_out/${TAG}_join $CMD > gpurun_out/${TAG}_join.log 2>&1
ls -la gpurun_out/
