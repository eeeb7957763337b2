"""Per-kernel timing (CUDA events, warm-up, rotating buffers larger than L2) -> gpurun_out/microbench.json.
Numbers are TFLOP/s or GB/s against MEASURED_PEAKS.json; used to steer optimisation, not as bench values."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3):
    """GPU time per call: `iters` calls captured into one CUDA graph (no host launch overhead), replayed 3x."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (3 * iters) * 1e-3


def r(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).half()


def main():
    res = []

    def rec(name, secs, flops=None, nbytes=None):
        d = {"name": name, "us": secs * 1e6}
        if flops:
            d["tflops"] = flops / secs / 1e12
        if nbytes:
            d["gbs"] = nbytes / secs / 1e9
        res.append(d)
        print(json.dumps(d), flush=True)

    for (M, N, K, geglu, tn) in [(2048, 10240, 1280, True, 256), (2048, 10240, 1280, True, 512),
                                 (2048, 1280, 5120, False, 192), (2048, 1280, 5120, False, 256), (2048, 1280, 5120, False, 0),
                                 (2048, 3840, 1280, False, 256), (2048, 3840, 1280, False, 512),
                                 (2048, 1280, 1280, False, 192), (2048, 1280, 1280, False, 256), (2048, 1280, 1280, False, 0),
                                 (8192, 5120, 640, True, 256), (8192, 5120, 640, True, 512),
                                 (8192, 640, 2560, False, 256), (8192, 640, 2560, False, 512),
                                 (8192, 1920, 640, False, 256), (8192, 1920, 640, False, 512),
                                 (8192, 640, 640, False, 128), (8192, 640, 640, False, 512),
                                 (16384, 10240, 1280, True, 512), (16384, 1280, 5120, False, 512),
                                 (8192, 8192, 8192, False, 256), (8192, 8192, 8192, False, 512)]:
        x, w, b = r(M, K), r(N, K, scale=K ** -0.5), r(N)
        try:
            t = timeit(lambda: ops.linear(x, w, b, geglu=geglu, tile_n=tn))
            rec(f"gemm M{M} N{N} K{K} geglu{int(geglu)} bn{tn}", t, 2.0 * M * N * K)
        except Exception as ex:  # noqa
            print("ERR", M, N, K, ex)

    for (B, H, Cin, Cout, s) in [(2, 32, 1280, 1280, 1), (2, 64, 640, 640, 1), (2, 128, 320, 320, 1),
                                 (2, 32, 2560, 1280, 1), (2, 64, 1920, 640, 1), (2, 128, 960, 320, 1),
                                 (2, 64, 1280, 1280, 1), (2, 128, 320, 320, 2), (16, 32, 1280, 1280, 1)]:
        x = r(B, H, H, Cin)
        w = r(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5)
        b = r(Cout)
        for tn in (192, 256, 0):
            try:
                t = timeit(lambda: ops.conv3x3(x, w, b, stride=s, tile_n=tn))
                rec(f"conv3x3 B{B} {H}^2 {Cin}->{Cout} s{s} bn{tn}", t, 2.0 * B * (H // s) ** 2 * Cout * 9 * Cin)
            except Exception as ex:  # noqa
                print("ERR conv", B, H, Cin, Cout, ex)

    for (B, H, N, Nk, nip) in [(2, 10, 4096, 4096, 0), (2, 20, 1024, 1024, 0), (16, 20, 1024, 1024, 0), (2, 20, 1024, 81, 4),
                               (2, 10, 4096, 77, 0), (16, 10, 4096, 4096, 0)]:
        q, k, v = r(B * N, H * 64), r(B * Nk, H * 64), r(B * Nk, H * 64)
        try:
            t = timeit(lambda: ops.attention(q, k, v, B, H, N, Nk, n_ip=nip, ip_scale=1.0), iters=10)
            rec(f"attn B{B} H{H} {N}x{Nk} ip{nip}", t, 4.0 * B * H * N * Nk * 64)
        except Exception as ex:  # noqa
            print("ERR attn", ex)

    for (B, H, C) in [(2, 128, 320), (2, 64, 640), (2, 32, 1280), (2, 32, 2560)]:
        x, g, bb = r(B, H, H, C), r(C), r(C)
        t = timeit(lambda: ops.groupnorm(x, g, bb, silu=True))
        rec(f"groupnorm B{B} {H}^2 C{C}", t, nbytes=3.0 * x.numel() * 2)
    for (rows, C) in [(2048, 1280), (8192, 640)]:
        x, g, bb = r(rows, C), r(C), r(C)
        t = timeit(lambda: ops.layernorm(x, g, bb))
        rec(f"layernorm {rows}x{C}", t, nbytes=2.0 * x.numel() * 2)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
