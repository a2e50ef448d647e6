#!/bin/bash
# GPU-side measurement recipe (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/run_gpu.sh r08 [tests] [bench] [launches] [full] [ref]'
# Everything lands in gpurun_out/<tag>_*; the summaries that are judged are copied into profiles/ by hand.
set -u
TAG=${1:-run}; shift || true
WHAT=" ${*:-tests bench launches full} "
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1

if [[ "$WHAT" == *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_tests.log
  tail -5 $OUT/${TAG}_tests.log
fi
if [[ "$WHAT" == *" bench "* ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  tail -c 3000 $OUT/${TAG}_bench.json
fi
if [[ "$WHAT" == *" ref "* ]]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
  tail -c 1500 $OUT/${TAG}_bench_ref.json
fi
if [[ "$WHAT" == *" launches "* ]]; then
  # launch list of the whole run (8 frames of ~210 launches); the summary keeps the whole frames between the first two L2 flushes
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-density --no-sweep --no-extra-warmup > $OUT/${TAG}_launches.log 2>&1
  python profiles/summarize_launches.py $OUT/${TAG}_launches.csv --frames > $OUT/${TAG}_launches.md 2>&1
  cat $OUT/${TAG}_launches.md
fi
if [[ "$WHAT" == *" full "* ]]; then
  for K in k_env_tc k_geom_tc k_shade_tc k_march_compact; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 140 --launch-count 1 -f \
        -o $OUT/${TAG}_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-density --no-sweep --no-extra-warmup > $OUT/${TAG}_full_$K.log 2>&1
    echo "ncu full $K exit $?"
  done
fi
