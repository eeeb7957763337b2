#!/bin/bash
IH_BENCH_SHAPES=1 python bench.py --steps 10 --warmup 3 --images 8 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/b8.err | tail -1 > gpurun_out/b8_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b8_bench.json').read())
print(d['value'], d['ms_per_step'], d['gpu_launches'])
for k,v in d['roofline']['families'].items(): print('  ',k, v['calls'], v['ms'], round(v.get('achieved',0),1), round(v.get('frac',0),3))
PY
head -30 gpurun_out/bench_shapes.md
