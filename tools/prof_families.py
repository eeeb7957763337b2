"""One launch of each kernel family at its most frequent SDXL shape (1024^2, UNet batch 2) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full ...` (profiles/README.md):
  plain GEMM 2048x1280x1280 (auto tile = 128x192, the <192,4,0,0> instantiation), FF-out 2048x1280x5120,
  implicit-GEMM conv3x3 B2 32^2 1280->1280, decoupled cross-attention B2 H20 N1024 L81 (attnx), self-attention
  B2 H20 N1024 (attn2 + combine), GroupNorm+SiLU B2 32^2 1280 (gn_stats + gn_apply)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402


def r(*s, scale=1.0):
    return (torch.randn(*s, device="cuda") * scale).half()


x, w, b = r(2048, 1280), r(1280, 1280, scale=1280 ** -0.5), r(1280)
res = r(2048, 1280)
st = torch.empty((20, 2048, 2), dtype=torch.float32, device="cuda")
x5, w5 = r(2048, 5120), r(1280, 5120, scale=5120 ** -0.5)
xc, wc, bc = r(2, 32, 32, 1280), r(1280, 9 * 1280, scale=(9 * 1280) ** -0.5), r(1280)
q = r(2048, 3840)
kv = r(2 * 81, 2560)
g, be = r(1280), r(1280)


def step():
    ops.linear(x, w, b, residual=res, stats_out=st)
    ops.linear(x5, w5, b, residual=res, stats_out=st)
    ops.conv3x3(xc, wc, bc)
    ops.attention(q[:, :1280], kv[:, :1280], kv[:, 1280:], 2, 20, 1024, 81, n_ip=4, ip_scale=1.0)
    ops.attention(q[:, :1280], q[:, 1280:2560], q[:, 2560:], 2, 20, 1024, 1024)
    ops.groupnorm(xc, g, be, groups=32, silu=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
