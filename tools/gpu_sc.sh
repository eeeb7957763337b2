#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -s 2>&1 | grep -E "passed|failed|rror|FAILED|assert" | tail -12
python -m pytest tests/test_unet_gpu.py -q -s -k "1024 or tiny or batch16" 2>&1 | grep -E "^\[|passed|failed|rror|FAILED|assert" | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/sc_bench.err | tail -1 > gpurun_out/sc_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/sc_bench.json').read())
print(d['value'], d['ms_per_step'], d['gpu_launches'])
for k,v in d['roofline']['families'].items(): print('  ',k, v['calls'], v['ms'], round(v.get('achieved',0),1), round(v.get('frac',0),3))
PY
