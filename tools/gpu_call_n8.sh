#!/bin/bash
# 8-GPU evidence: bench contract line (PNS tail + pns block with the CLIP judge), C4 PNS N=32 x 50 steps (CLIP judge),
# two-phase PNS; `SWEEP=1` adds the C5 sweep rows (tools/sweep.py under torchrun).
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/n8_bench.err | grep '^{' | tail -1 > gpurun_out/n8_bench.json
$TR --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 3 --pns 32 2>gpurun_out/n8_pns.err | grep '^{' | tail -1 > gpurun_out/n8_pns32.json
$TR --master-port 29523 bench.py --gpus 8 --steps 50 --warmup 3 --pns 32 --pns-preview 10 2>/dev/null | grep '^{' | tail -1 > gpurun_out/n8_pns32_two_phase.json
if [ "${SWEEP:-0}" = "1" ]; then IH_SWEEP_STEPS=20 $TR --master-port 29524 tools/sweep.py > gpurun_out/n8_sweep.log 2>&1; fi
cut -c1-300 gpurun_out/n8_bench.json; echo; cut -c1-800 gpurun_out/n8_pns32.json; echo; cut -c1-500 gpurun_out/n8_pns32_two_phase.json
