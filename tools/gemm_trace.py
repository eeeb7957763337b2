"""Ramp breakdown of one GEMM launch inside a dependent chain (globaltimer stamps of CTA 0)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
names = ["entry", "setup done", "pdl_wait done", "first operands", "acc0 ready", "acc1 ready", "acc2 ready", "acc3+ ready", "cta done",
         "slab0 in regs", "slab0 in smem", "slab1/mid in smem", "last slab in smem", "stores read out"]


def run(label, M, N, K, geglu=False, tn=0, full=False):
    x = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    g = torch.ones(K, device="cuda").half()
    kw = {}
    if full:
        kw = dict(bias=torch.randn(N, device="cuda").half(), residual=torch.randn(M, N, device="cuda").half(),
                  stats_out=torch.empty((N // 64, M, 2), dtype=torch.float32, device="cuda"))
    for _ in range(3):
        n = ops.layernorm(x, g, g)
        ops.linear(n, w, geglu=geglu, tile_n=tn, **kw)
    torch.cuda.synchronize()
    lib.ih_gemm_set_trace(buf.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = ops.layernorm(x, g, g)
    s.record()
    ops.linear(n, w, geglu=geglu, tile_n=tn, **kw)
    e.record()
    torch.cuda.synchronize()
    lib.ih_gemm_set_trace(None)
    t = buf.cpu().tolist()
    rel = [(t[i] - t[0]) / 1e3 if t[i] else None for i in range(14)]
    print(label, f"event {s.elapsed_time(e) * 1e3:.1f} us |", ", ".join(f"{n}={v:.2f}" for n, v in zip(names, rel) if v is not None))
    buf.zero_()


run("1 tile 128x256x64      ", 128, 256, 64, tn=256)
run("1280^2 (80 tiles)      ", 2048, 1280, 1280, tn=256)
run("1280^2 auto (112 tiles) ", 2048, 1280, 1280, tn=0)
run("1280^2 auto +bias+res+stats", 2048, 1280, 1280, tn=0, full=True)
run("1280^2 bn128 (160 tiles)", 2048, 1280, 1280, tn=128)
run("QKV 2048x3840x1280     ", 2048, 3840, 1280, tn=256)
run("FF-in geglu            ", 2048, 10240, 1280, geglu=True, tn=256)
run("FF-out 2048x1280x5120  ", 2048, 1280, 5120, tn=256)
run("FF-in geglu pair       ", 2048, 10240, 1280, geglu=True, tn=512)
