"""Fixed per-launch cost of the kernels inside a CUDA graph: tiny problems, 50 back-to-back dependent launches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def r(*s):
    return torch.randn(*s, device="cuda").half()


x, w = r(128, 64), r(256, 64)
print("gemm 1 tile bn256 (us):", timeit(lambda: ops.linear(x, w, tile_n=256), iters=50) * 1e6)
print("gemm 1 tile pair  (us):", timeit(lambda: ops.linear(r(256, 64), w, tile_n=512), iters=50) * 1e6)
xl, g, b = r(256, 1280), r(1280), r(1280)
print("layernorm 256x1280 (us):", timeit(lambda: ops.layernorm(xl, g, b), iters=50) * 1e6)
# alternating LN -> GEMM chain like a transformer block
xa, wa = r(2048, 1280), r(1280, 1280)
def chain():
    n = ops.layernorm(xa, g, b)
    ops.linear(n, wa)
print("LN+GEMM(2048x1280x1280) pair-auto (us):", timeit(chain, iters=25) * 1e6)
print("GEMM alone (us):", timeit(lambda: ops.linear(xa, wa), iters=50) * 1e6)
print("LN alone 2048x1280 (us):", timeit(lambda: ops.layernorm(xa, g, b), iters=50) * 1e6)
q = r(2048, 3840)
print("attn 2x20x1024 (us):", timeit(lambda: ops.attention(q[:, :1280], q[:, 1280:2560], q[:, 2560:], 2, 20, 1024, 1024), iters=25) * 1e6)
