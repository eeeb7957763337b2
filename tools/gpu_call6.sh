#!/bin/bash
# 2-GPU check of the bench contract: NCCL ranks, PNS tail in the timed region, pns block with the CLIP judge.
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
tail -c 2500 gpurun_out/c6_bench_n2.json; tail -5 gpurun_out/c6_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
