mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_field.py -x -q -m gpu 2>&1 | tail -4
timeout 100 python profiles/env_timeline.py > gpurun_out/r17_env_timeline.txt 2>&1; cut -c1-330 gpurun_out/r17_env_timeline.txt | sed -n 2,5p
timeout 400 python bench.py --steps 10 --warmup 3 --no-train --no-gpu-reference --no-cpu-baseline > gpurun_out/r17_bench.json 2>gpurun_out/r17_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r17_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms_per_step'],d['roofline']['frac'])"
ENVIDR_ENV_TC_CTAS=3 timeout 400 python bench.py --steps 10 --warmup 3 --no-train --no-gpu-reference --no-cpu-baseline > gpurun_out/r17_bench_mc.json 2>gpurun_out/r17_bench_mc.err; python -c "
import json;d=json.load(open('gpurun_out/r17_bench_mc.json'));print('MC',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms_per_step'],d['roofline']['frac'])"
