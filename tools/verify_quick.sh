#!/bin/bash
# Quick GPU verification after a kernel change: kernel parity tests, the SDXL-size UNet parity tests, then a same-box
# environment-switch A/B of the step time (tools/ab_env.sh; pass the settings to compare in AB_CONFIGS).
set -u
echo "kernels: $(python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -1)"
python -m pytest tests/test_unet_gpu.py -q -s -k "512 or 1024 or tiny" 2>&1 | grep -E "^\[unet SDXL|passed|failed|FAILED|Error" | tail -6
bash tools/ab_env.sh
