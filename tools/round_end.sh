#!/bin/bash
# Round-end evidence run on one B200 (see profiles/README.md): tests, smoke, both bench arms, launch list, ncu --set full
# per kernel family, GEMM timeline trace, resolution x batch sweep, micro-benchmarks.  Outputs -> gpurun_out/re_*.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|rror|FAILED" > gpurun_out/re_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/re_smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 2>gpurun_out/re_bench.err | tail -1 > gpurun_out/re_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/re_ref.err | tail -1 > gpurun_out/re_bench_reference.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/re_launches.csv python tools/profile_step.py > gpurun_out/re_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/re_launches.csv > gpurun_out/re_launches.md
ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/re_families -f \
    python tools/prof_families.py > gpurun_out/re_ncu_families.log 2>&1
python tools/gemm_trace.py > gpurun_out/re_gemm_trace.txt 2>&1
python tools/sweep.py > gpurun_out/re_sweep.log 2>&1
python tools/microbench.py > gpurun_out/re_microbench.log 2>&1
python tools/edit_latency.py > gpurun_out/re_edit_latency.log 2>&1
python tools/layer_head_to_head.py > gpurun_out/re_layer.log 2>&1
IH_XATTN_FUSED=1 python tools/layer_head_to_head.py > gpurun_out/re_layer_fused.log 2>&1
python tools/eager_ref_gpu.py 1024 1 > gpurun_out/re_eager.log 2>&1
tail -3 gpurun_out/re_pytest.txt; tail -2 gpurun_out/re_smoke.txt; cut -c1-400 gpurun_out/re_bench_n1.json; echo; cut -c1-300 gpurun_out/re_bench_reference.json; echo
tail -20 gpurun_out/re_sweep.log
