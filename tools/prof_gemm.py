"""A few big GEMM launches for `ncu --set full` (tile variants 256 = single CTA 128x256, 512 = CTA pair 256x256)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402

M, N, K = 8192, 8192, 8192
x = (torch.randn(M, K, device="cuda")).half()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
for tn in (256, 512):
    for _ in range(2):
        ops.linear(x, w, tile_n=tn)
x2 = torch.randn(2048, 1280, device="cuda").half()
w2 = (torch.randn(10240, 1280, device="cuda") * 1280 ** -0.5).half()
b2 = torch.randn(10240, device="cuda").half()
for tn in (256, 512):
    for _ in range(2):
        ops.linear(x2, w2, b2, geglu=True, tile_n=tn)
torch.cuda.synchronize()
