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


for (M, N, K) in [(2048, 1280, 1280), (2048, 1280, 5120), (2048, 3840, 1280), (8192, 640, 2560), (8192, 640, 640),
                  (16384, 1280, 1280)]:
    x, w, b, res = r(M, K), r(N, K, scale=K ** -0.5), r(N), r(M, N)
    st = torch.empty((N // 64, M, 2), dtype=torch.float32, device="cuda")
    for full in (False, True):
        row = {"M": M, "N": N, "K": K, "epilogue": "bias+res+stats" if full else "plain"}
        for tn in (0, 128, 192, 256, 384, 512):
            kw = dict(bias=b, residual=res, stats_out=st) if full else {}
            try:
                t = timeit(lambda: ops.linear(x, w, tile_n=tn, **kw), iters=10)
                row[f"bn{tn}_us"] = round(t * 1e6, 1)
            except Exception as ex:  # noqa: BLE001
                row[f"bn{tn}_us"] = str(ex)[:40]
        row["best_tflops"] = round(2.0 * M * N * K / (min(v for k, v in row.items() if k.endswith("_us") and isinstance(v, float)) * 1e-6) / 1e12)
        print(json.dumps(row), flush=True)


# implicit-GEMM convs at UNet batch 2 (1024^2): tile variants incl. the CTA-pair 256x192 tile (tile_n = 384)
for (B, H, Cin, Cout) in [(2, 32, 1280, 1280), (2, 64, 640, 640), (2, 128, 320, 320), (2, 32, 2560, 1280), (2, 64, 1920, 640)]:
    x, w, b = r(B, H, H, Cin), r(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5), r(Cout)
    row = {"conv": f"B{B} {H}^2 {Cin}->{Cout}"}
    for tn in (0, 192, 256, 384, 512):
        try:
            t = timeit(lambda: ops.conv3x3(x, w, b, tile_n=tn), iters=10)
            row[f"bn{tn}_us"] = round(t * 1e6, 1)
        except Exception as ex:  # noqa: BLE001
            row[f"bn{tn}_us"] = str(ex)[:40]
    print(json.dumps(row), flush=True)
