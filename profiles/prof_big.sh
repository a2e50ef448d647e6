set -u
OUT=gpurun_out; mkdir -p $OUT
for spec in k_geom_tc:2 k_shade_tc:1 k_march_compact:1 k_env_tc:1 k_composite_compact:1; do
  K=${spec%%:*}; S=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $S --launch-count 1 -f \
      -o $OUT/r08b_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train --no-density --no-extra-warmup > $OUT/r08b_full_$K.log 2>&1
  echo "ncu full $K exit $?"
done
