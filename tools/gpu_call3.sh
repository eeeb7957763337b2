#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -s 2>&1 | grep -E "passed|failed|rror|FAILED|assert" | tail -15 > gpurun_out/c3_pytest_kernels.txt
python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py -q -s 2>&1 | grep -E "^\[|passed|failed|rror|FAILED|assert" | tail -60 > gpurun_out/c3_pytest_rest.txt
python tools/gemm_trace.py > gpurun_out/c3_gemm_trace.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>gpurun_out/c3_bench.err | tail -1 > gpurun_out/c3_bench_n1.json
IH_GN_FUSED=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-families 2>/dev/null | tail -1 > gpurun_out/c3_bench_gn2k.json
cat gpurun_out/c3_pytest_kernels.txt; tail -30 gpurun_out/c3_pytest_rest.txt; cat gpurun_out/c3_gemm_trace.txt; cut -c1-300 gpurun_out/c3_bench_n1.json; echo; cut -c1-300 gpurun_out/c3_bench_gn2k.json
