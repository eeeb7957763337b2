#!/bin/bash
# Final evidence for the committed tree on one B200: full GPU suite, smoke, N=1 bench line, launch list of one step.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|rror|FAILED" > gpurun_out/fe_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fe_smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 2>gpurun_out/fe_bench.err | tail -1 > gpurun_out/fe_bench_n1.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/fe_launches.csv python tools/profile_step.py > gpurun_out/fe_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/fe_launches.csv > gpurun_out/fe_launches.md
tail -2 gpurun_out/fe_pytest.txt; tail -2 gpurun_out/fe_smoke.txt; cut -c1-300 gpurun_out/fe_bench_n1.json; echo; head -12 gpurun_out/fe_launches.md
