#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -s 2>&1 | grep -E "passed|failed|rror|FAILED|assert" | tail -15 > gpurun_out/c7_pytest_kernels.txt
python tools/gemm_trace.py > gpurun_out/c7_gemm_trace.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/c7_bench.err | tail -1 > gpurun_out/c7_bench_n1.json
python -m pytest tests/test_unet_gpu.py tests/test_clip_gpu.py -q -s -k "1024 or tiny or goldens or mini or scorer" 2>&1 | grep -E "^\[|passed|failed|rror|FAILED|assert" | tail -20 > gpurun_out/c7_pytest_rest.txt
cat gpurun_out/c7_pytest_kernels.txt; tail -12 gpurun_out/c7_pytest_rest.txt; cat gpurun_out/c7_gemm_trace.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c7_bench_n1.json').read())
print(d['value'], d['ms_per_step'], d['gpu_launches'])
for k,v in d['roofline']['families'].items(): print('  ',k, v['calls'], v['ms'], round(v.get('achieved',0),1), round(v.get('frac',0),3))
PY
