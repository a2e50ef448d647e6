#!/bin/bash
# Round-2 (third session) evidence run:  gpurun --timeout 2400 -- 'bash profiles/run_final_r3.sh'
set -u
OUT=gpurun_out; TAG=${1:-r3g}; mkdir -p $OUT
NOX="--no-cpu-baseline --no-gpu-reference --no-train --no-density --no-sweep --no-extra-warmup"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_tests.log; tail -3 $OUT/${TAG}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 600 $OUT/${TAG}_bench_ref.json
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1200 $OUT/${TAG}_bench.json
timeout 600 python bench.py --config neus --steps 3 > $OUT/${TAG}_bench_neus.json 2> $OUT/${TAG}_bench_neus.err; tail -c 600 $OUT/${TAG}_bench_neus.json
timeout 120 python profiles/env_timeline.py > $OUT/${TAG}_env_timeline.txt 2>&1
timeout 300 python profiles/train_profile.py > $OUT/${TAG}_train_profile.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 $NOX > $OUT/${TAG}_launches.log 2>&1
python profiles/summarize_launches.py $OUT/${TAG}_launches.csv --frames > $OUT/${TAG}_launches.md 2>&1; head -14 $OUT/${TAG}_launches.md
for KS in k_env_tc:5 k_geom_tc:30 k_shade_tc:5; do
  K=${KS%%:*}; S=${KS##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $S --launch-count 1 -f -o $OUT/${TAG}_$K python bench.py --steps 1 --warmup 3 $NOX > $OUT/${TAG}_full_$K.log 2>&1
  echo "ncu full $K exit $?"
done
# training kernels: profiles/train_profile.py launches only the SAVE instantiation of k_env_tc and the three modes of k_chain_tc
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_env_tc --launch-skip 4 --launch-count 1 -f -o $OUT/${TAG}_train_k_env_tc_save python profiles/train_profile.py > $OUT/${TAG}_full_train_k_env_tc_save.log 2>&1
echo "ncu full k_env_tc<SAVE> exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_chain_tc<0>" --launch-skip 4 --launch-count 1 -f -o $OUT/${TAG}_train_k_chain_tc_0 python profiles/train_profile.py > $OUT/${TAG}_full_train_k_chain_tc_0.log 2>&1
echo "ncu full k_chain_tc<0> exit $?"
