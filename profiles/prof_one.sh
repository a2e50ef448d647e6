#!/bin/bash
# ncu --set full of one launch of one kernel:  bash profiles/prof_one.sh <tag> <kernel-regex> <launch-skip>
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 -f \
    -o $OUT/$1_$2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-density --no-sweep --no-extra-warmup > $OUT/$1_full_$2.log 2>&1
echo "ncu full $2 exit $?"
