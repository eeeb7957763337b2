"""Time the oracle UNet (the reference's processors + the restated diffusers UNet) in torch-eager fp16 on one GPU:
the stand-in for "the reference GPU diffusers path" (cuDNN / cuBLAS / torch SDPA as torch dispatches them).
Writes gpurun_out/eager_ref.json.  Not a bench arm; context for DESIGN.md / profiles."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200.config import SDXL_BASE as cfg  # noqa: E402
from oracle import adapter_ref as A  # noqa: E402
from oracle.unet_ref import UNetRef  # noqa: E402


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lat = res // 8
    torch.backends.cuda.matmul.allow_tf32 = True
    with torch.device("meta"):
        m = UNetRef(cfg)
    m = m.to_empty(device="cuda").half()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.fill_(0.5)
            else:
                p.uniform_(-0.02, 0.02)
    with torch.device("cuda"):
        A.install_processors(m, cfg, dtype=torch.float16)
    m.eval()
    B = 2 * n
    x = torch.randn(B, 4, lat, lat, device="cuda").half()
    ehs = torch.randn(B, 81, 2048, device="cuda").half()
    te = torch.randn(B, 1280, device="cuda").half()
    tid = torch.tensor([[res, res, 0, 0, res, res]] * B, device="cuda", dtype=torch.float32)
    out = {}
    with torch.no_grad():
        for _ in range(3):
            m(x, 500.0, ehs, te, tid)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        K = 10
        for _ in range(K):
            m(x, 500.0, ehs, te, tid)
        e.record()
        torch.cuda.synchronize()
        out["eager_ms_per_unet_forward"] = s.elapsed_time(e) / K
        print(json.dumps(out), flush=True)
        # same thing replayed as a CUDA graph (removes the eager launch overhead: the best case for library kernels)
        tt = torch.full((B,), 500.0, device="cuda")
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            m(x, tt, ehs, te, tid)
        torch.cuda.current_stream().wait_stream(st)
        with torch.cuda.graph(g):
            y = m(x, tt, ehs, te, tid)
        g.replay()
        torch.cuda.synchronize()
        s.record()
        for _ in range(K):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        out["graph_ms_per_unet_forward"] = s.elapsed_time(e) / K
    out.update({"res": res, "images": n, "unet_batch": B, "torch": torch.__version__})
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"eager_ref_{res}_{n}.json"), "w"))


if __name__ == "__main__":
    main()
