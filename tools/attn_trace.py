"""Where does a KV iteration of the ping-pong attention kernel spend its time?  clock64 stamps of CTA 0, blocks 4..7.
Needs a library built with -DIH_ATTN_TRACE=1 (make -C imagharmony_b200/csrc EXTRA=-DIH_ATTN_TRACE=1)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200 import _lib, ops  # noqa: E402


def main():
    B, H, N = 2, 10, 4096
    C = H * 64
    qkv = (torch.randn(B * N, 3 * C, device="cuda")).half()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    lib = _lib.load()
    buf = torch.zeros(192, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.attention(q, k, v, B, H, N, N, kv_split=False)
    lib.ih_attention_set_trace(buf.data_ptr())
    ops.attention(q, k, v, B, H, N, N, kv_split=False)
    torch.cuda.synchronize()
    lib.ih_attention_set_trace(None)
    t = buf.cpu().tolist()
    t0 = t[0]
    names = ["wait S", "S seen", "S in regs", "max done", "pair barrier", "exp+P done", "arrived"]
    for tile in range(2):
        for j in range(4):
            row = t[tile * 64 + j * 8: tile * 64 + j * 8 + 7]
            print(f"tile {tile} block {4 + j}: " + "  ".join(f"{n} {x - t0:6d}" for n, x in zip(names, row)))


if __name__ == "__main__":
    main()
