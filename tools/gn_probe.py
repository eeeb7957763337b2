import os, sys, torch
sys.path.insert(0, os.getcwd())
from imagharmony_b200 import ops
from tools.microbench import timeit, r
for (B, H, C) in [(2, 128, 320), (2, 64, 640), (2, 32, 1280), (2, 32, 2560), (2, 64, 1920), (16, 32, 1280)]:
    x = r(B, H, H, C); g = r(C); b = r(C)
    t = timeit(lambda: ops.groupnorm(x, g, b, silu=True))
    print(f"groupnorm B{B} {H}^2 C{C}: {t*1e6:.1f} us  ({3*x.numel()*2/t/1e9:.0f} GB/s)", flush=True)
