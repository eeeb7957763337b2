#!/bin/bash
set -u
mkdir -p gpurun_out
IH_BENCH_SHAPES=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/c9_bench.err | tail -1 > gpurun_out/c9_bench_n1.json
IH_GEGLU_PAIR=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 > gpurun_out/c9_bench_gpair.json
python tools/edit_latency.py > gpurun_out/c9_edit_latency.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/c9_bench_n1.json','gpurun_out/c9_bench_gpair.json'):
    d=json.loads(open(f).read())
    print(f, d['value'], d['ms_per_step'], d['gpu_launches'])
    for k,v in d['roofline']['families'].items(): print('  ',k, v['calls'], v['ms'], round(v.get('achieved',0),1), round(v.get('frac',0),3))
PY
head -40 gpurun_out/bench_shapes.md; tail -3 gpurun_out/c9_edit_latency.log
