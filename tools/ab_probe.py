import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from imagharmony_b200 import ops
from tools.microbench import timeit, r
res = {}
for (M, N, K, geglu, tn) in [(2048, 1280, 1280, False, 0), (2048, 3840, 1280, False, 0), (8192, 1920, 640, False, 0), (2048, 10240, 1280, True, 0), (2048, 1280, 5120, False, 0)]:
    x, w, b = r(M, K), r(N, K, scale=K ** -0.5), r(N)
    t = timeit(lambda: ops.linear(x, w, b, geglu=geglu, tile_n=tn))
    print(f"gemm M{M} N{N} K{K} g{int(geglu)}: {t*1e6:.1f} us", flush=True)
for (B, H, Cin, Cout, s) in [(2, 32, 1280, 1280, 1), (2, 64, 640, 640, 1), (2, 128, 320, 320, 1), (2, 32, 2560, 1280, 1), (2, 64, 1920, 640, 1)]:
    x = r(B, H, H, Cin); w = r(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5); b = r(Cout)
    for tn in (192, 256, 0):
        t = timeit(lambda: ops.conv3x3(x, w, b, stride=s, tile_n=tn))
        print(f"conv B{B} {H}^2 {Cin}->{Cout} bn{tn}: {t*1e6:.1f} us", flush=True)
