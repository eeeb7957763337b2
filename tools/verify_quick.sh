#!/bin/bash
set -u
echo "kernels: $(IH_BN64_CONV=1 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -1)"
IH_BN64_CONV=1 python -m pytest tests/test_unet_gpu.py -q -s -k "512 or tiny" 2>&1 | grep -E "^\[unet SDXL|passed|failed|FAILED|Error" | tail -6
AB_CONFIGS='BASE=1 IH_BN64_CONV=1' bash tools/ab_env.sh
