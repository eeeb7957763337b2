"""Small A/B probes (graph-timed, same process => same thermal state).  `python tools/ab_probe.py tile` measures the
main-loop cost per 64-deep k-block of the 128xBN tiles with 144 of 148 SMs busy (calibrates pick_bn in gemm.cu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402
from tools.microbench import timeit, r  # noqa: E402


def tile_costs():
    M = 2048
    for bn in (128, 192, 256):
        N = 9 * bn                      # 9 x 16 = 144 tiles: one wave
        ts = []
        for K in (1280, 5120):
            x, w = r(M, K), r(N, K, scale=K ** -0.5)
            ts.append(timeit(lambda: ops.linear(x, w, tile_n=bn)))
        per_kb = (ts[1] - ts[0]) / ((5120 - 1280) / 64)
        print(f"128x{bn}: K=1280 {ts[0]*1e6:.1f} us, K=5120 {ts[1]*1e6:.1f} us -> {per_kb*1e6:.3f} us per k-block, "
              f"fixed {(ts[0] - 20 * per_kb)*1e6:.1f} us", flush=True)


def shapes():
    for (M, N, K, geglu, tn) in [(2048, 1280, 1280, False, 0), (2048, 3840, 1280, False, 0), (8192, 1920, 640, False, 0),
                                 (2048, 10240, 1280, True, 0), (2048, 1280, 5120, False, 0)]:
        x, w, b = r(M, K), r(N, K, scale=K ** -0.5), r(N)
        t = timeit(lambda: ops.linear(x, w, b, geglu=geglu, tile_n=tn))
        print(f"gemm M{M} N{N} K{K} g{int(geglu)}: {t*1e6:.1f} us", flush=True)
    for (B, H, Cin, Cout, s) in [(2, 32, 1280, 1280, 1), (2, 64, 640, 640, 1), (2, 128, 320, 320, 1), (2, 32, 2560, 1280, 1),
                                 (2, 64, 1920, 640, 1)]:
        x = r(B, H, H, Cin)
        w = r(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5)
        b = r(Cout)
        for tn in (128, 192, 256, 0):
            t = timeit(lambda: ops.conv3x3(x, w, b, stride=s, tile_n=tn))
            print(f"conv B{B} {H}^2 {Cin}->{Cout} bn{tn}: {t*1e6:.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tile":
        tile_costs()
    else:
        shapes()
