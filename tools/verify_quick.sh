#!/bin/bash
set -u
python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -2
python -m pytest tests/test_unet_gpu.py -q -s -k "1024" 2>&1 | grep -E "^\[unet SDXL|passed|failed|FAILED" | tail -6
bash tools/ab_env.sh
