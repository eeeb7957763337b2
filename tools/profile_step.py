"""One eager (non-graph) denoise step of the full SDXL configuration between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (see profiles/README.md for the exact command lines)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from imagharmony_b200.config import SDXL_BASE as cfg  # noqa: E402
from imagharmony_b200.denoise import DenoiseEngine  # noqa: E402


def main():
    res = int(os.environ.get("IH_RES", "1024"))
    n = int(os.environ.get("IH_IMAGES", "1"))
    lat = res // 8
    torch.cuda.set_device(0)
    unet = bench.build_native(cfg, torch.device("cuda", 0))
    eng = DenoiseEngine(unet, use_cuda_graph=False)
    latents, pos, neg, pooled, npooled, tid = bench.synth_inputs(cfg, n, lat, 50, 0)
    eng.run(latents, pos, neg, pooled, npooled, tid, 50, stop_after=1)       # warm-up (caches, attributes)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.run(latents, pos, neg, pooled, npooled, tid, 50, stop_after=1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
