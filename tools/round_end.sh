#!/bin/bash
# Round-end evidence run on one B200 (see profiles/README.md): tests, smoke, bench arms, launch list, ncu --set full
# captures of the dominant GEMM and the ping-pong attention kernel, resolution x batch sweep.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/re_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/re_smoke.txt 2>&1
python bench.py 2>gpurun_out/re_bench.err | tail -1 > gpurun_out/re_bench_n1.json
python bench.py --impl reference 2>gpurun_out/re_ref.err | tail -1 > gpurun_out/re_bench_reference.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/re_launches.csv python tools/profile_step.py > gpurun_out/re_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/re_launches.csv > gpurun_out/re_launches.md
ncu --set full --clock-control none --import-source on -k regex:gemm_f16_kernel -s 4 -c 1 -o gpurun_out/re_geglu256 \
    -f python tools/prof_gemm.py > gpurun_out/re_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn2_f16_kernel -c 2 \
    -o gpurun_out/re_attn2 -f python tools/profile_step.py > gpurun_out/re_ncu2.log 2>&1
python tools/sweep.py > gpurun_out/re_sweep.log 2>&1
python tools/microbench.py > gpurun_out/re_microbench.log 2>&1
cat gpurun_out/re_pytest.txt gpurun_out/re_smoke.txt; cut -c1-400 gpurun_out/re_bench_n1.json; cut -c1-300 gpurun_out/re_bench_reference.json
tail -12 gpurun_out/re_sweep.log
