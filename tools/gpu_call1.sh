#!/bin/bash
# Round-2 GPU call 1: full GPU test-suite, smoke, both bench arms, launch list of one step, ncu --set full per family.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "^\[|passed|failed|error|Error|FAILED|assert" | tail -80 > gpurun_out/r2_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_bench.err | tail -1 > gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/r2_ref.err | tail -1 > gpurun_out/r2_bench_reference.json
python tools/gemm_trace.py > gpurun_out/r2_gemm_trace.txt 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_families -f \
    python tools/prof_families.py > gpurun_out/r2_ncu_families.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py > gpurun_out/r2_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches.md
cat gpurun_out/r2_pytest.txt | tail -30; tail -3 gpurun_out/r2_smoke.txt; cut -c1-600 gpurun_out/r2_bench_n1.json; cut -c1-300 gpurun_out/r2_bench_reference.json
