"""A/B of the GEMM tile variants for the residual-stream GEMMs (bias + residual + row statistics) at several M:
tile_n 192 / 256 = single-CTA 128 x BN, 512 = CTA pair 256 x 256, 0 = the library's choice.  Graph-timed (tools/microbench.timeit)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def r(*s, scale=1.0):
    return (torch.randn(*s, device="cuda") * scale).half()


for (M, N, K) in [(16384, 1280, 1280), (8192, 1280, 1280), (4096, 1280, 1280), (2048, 1280, 1280), (16384, 3840, 1280),
                  (65536, 640, 640), (32768, 640, 640)]:
    x, w, b, res = r(M, K), r(N, K, scale=K ** -0.5), r(N), r(M, N)
    st = torch.empty((N // 64, M, 2), dtype=torch.float32, device="cuda")
    for full in (False, True):
        row = {"M": M, "N": N, "K": K, "epilogue": "bias+res+stats" if full else "plain"}
        for tn in (0, 128, 192, 256, 512):
            kw = dict(bias=b, residual=res, stats_out=st) if full else {}
            try:
                t = timeit(lambda: ops.linear(x, w, tile_n=tn, **kw), iters=10)
                row[f"bn{tn}_us"] = round(t * 1e6, 1)
            except Exception as ex:  # noqa: BLE001
                row[f"bn{tn}_us"] = str(ex)[:40]
        row["best_tflops"] = round(2.0 * M * N * K / (min(v for k, v in row.items() if k.endswith("_us") and isinstance(v, float)) * 1e-6) / 1e12)
        print(json.dumps(row), flush=True)
