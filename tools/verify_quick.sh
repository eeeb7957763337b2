#!/bin/bash
set -u
for v in 1 2 3; do
  echo "variant $v: $(IH_ATTN_VARIANT=$v python -m pytest tests/test_kernels_gpu.py -q -k 'attention or attn' 2>&1 | tail -1)"
done
AB_CONFIGS='BASE=1 IH_ATTN_VARIANT=1 IH_ATTN_VARIANT=2 IH_ATTN_VARIANT=3' bash tools/ab_env.sh
