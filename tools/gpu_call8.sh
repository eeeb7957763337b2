#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -s -k "gemm or conv" 2>&1 | grep -E "passed|failed|rror|FAILED|assert" | tail -8 > gpurun_out/c8_pytest_kernels.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/c8_bench.err | tail -1 > gpurun_out/c8_bench_n1.json
IH_PREFETCH=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/c8_bench_nopf.json
python -m pytest tests/test_unet_gpu.py -q -s -k "1024 or tiny" 2>&1 | grep -E "^\[|passed|failed|rror|FAILED|assert" | tail -12 > gpurun_out/c8_pytest_rest.txt
cat gpurun_out/c8_pytest_kernels.txt; tail -8 gpurun_out/c8_pytest_rest.txt
python - <<'PY'
import json
for f in ('gpurun_out/c8_bench_n1.json','gpurun_out/c8_bench_nopf.json'):
    d=json.loads(open(f).read())
    print(f, d['value'], d['ms_per_step'], d['gpu_launches'])
    for k,v in d['roofline']['families'].items(): print('  ',k, v['calls'], v['ms'], round(v.get('achieved',0),1), round(v.get('frac',0),3))
PY
