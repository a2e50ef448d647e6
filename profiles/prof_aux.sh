#!/bin/bash
# ncu --set full of the round-1 "next row" kernels, one launch each:  bash profiles/prof_aux.sh <tag>
set -u
TAG=${1:-aux}; OUT=gpurun_out; mkdir -p $OUT
cap() {   # cap <kernel regex> <launch-skip> <script...>
  local K=$1 SKIP=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP --launch-count 1 -f \
      -o $OUT/${TAG}_$K "$@" > $OUT/${TAG}_full_$K.log 2>&1
  echo "ncu full $K exit $?"
}
cap k_adam 3 python profiles/optim_bench.py
cap k_density_ema 3 python profiles/density_bench.py
cap k_density_points 3 python profiles/density_bench.py
cap k_density_pack 3 python profiles/density_bench.py
cap k_loss_fwd 3 python profiles/epilogue_bench.py
cap k_loss_bwd 3 python profiles/epilogue_bench.py
cap k_get_rays 3 python profiles/epilogue_bench.py
cap k_pow2_scales 30 python profiles/train_profile.py
