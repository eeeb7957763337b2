"""A/B timing of the LayerNorm fold: (layernorm kernel + plain GEMM) vs (GEMM with folded LN), and the cost of writing
row statistics in the producer.  Graph-timed like tools/microbench.py.  Steering only, not bench values."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import ops  # noqa: E402
from tools.microbench import timeit, r  # noqa: E402


def main():
    for (M, N, K, geglu) in [(2048, 3840, 1280, False), (2048, 1280, 1280, False), (2048, 10240, 1280, True),
                             (8192, 1920, 640, False), (8192, 640, 640, False), (8192, 5120, 640, True)]:
        x = r(M, K)
        w, b = r(N, K, scale=K ** -0.5), r(N)
        gamma, beta = (1 + 0.1 * r(K).float()).half(), (0.1 * r(K).float()).half()
        w_c, c = ops.fold_layernorm(w, b, gamma, beta)
        slabs = K // 64
        st = torch.empty((slabs, M, 2), dtype=torch.float32, device="cuda")
        # a producer with the consumer's K as its N
        wp, bp = r(K, K, scale=K ** -0.5), r(K)
        res = r(M, K)
        hp = torch.empty(M, K, dtype=torch.float16, device="cuda")
        t_p0 = timeit(lambda: ops.linear(x, wp, bp, residual=res, out=hp))
        t_p1 = timeit(lambda: ops.linear(x, wp, bp, residual=res, out=hp, stats_out=st))
        t_ln = timeit(lambda: ops.layernorm(hp, gamma, beta, 1e-5))
        t_g0 = timeit(lambda: ops.linear(hp, w, b, geglu=geglu))
        t_g1 = timeit(lambda: ops.linear(hp, w_c, c, geglu=geglu, ln=(st, 1e-5)))

        def seq0():
            ops.linear(x, wp, bp, residual=res, out=hp)
            n = ops.layernorm(hp, gamma, beta, 1e-5)
            ops.linear(n, w, b, geglu=geglu)

        def seq1():
            ops.linear(x, wp, bp, residual=res, out=hp, stats_out=st)
            ops.linear(hp, w_c, c, geglu=geglu, ln=(st, 1e-5))
        t_s0, t_s1 = timeit(seq0), timeit(seq1)
        print(f"M{M} N{N} K{K} geglu{int(geglu)}: producer {t_p0*1e6:.1f} -> +stats {t_p1*1e6:.1f} us | layernorm {t_ln*1e6:.1f} us | "
              f"consumer {t_g0*1e6:.1f} -> folded {t_g1*1e6:.1f} us | sequence {t_s0*1e6:.1f} -> {t_s1*1e6:.1f} us", flush=True)


if __name__ == "__main__":
    main()
