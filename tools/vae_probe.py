"""Time the native VAE decode (scope row f1) at 1024x1024 (latent 128x128) with random-init SDXL-VAE weights, and the
oracle decoder in torch-eager fp16 / fp32 on the same GPU for context.  Writes gpurun_out/vae_probe.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200.config import SDXL_VAE as cfg  # noqa: E402
from imagharmony_b200.vae import AutoencoderKLDecoder  # noqa: E402
from imagharmony_b200.weights import random_state_dict, shapes_of  # noqa: E402
from oracle.vae_ref import VAEDecoderRef  # noqa: E402


def ev_time(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    with torch.device("meta"):
        shapes = shapes_of(VAEDecoderRef(cfg))
    sd = random_state_dict(shapes, 0)
    native = AutoencoderKLDecoder.from_state_dict(cfg, sd, device="cuda")
    ref = VAEDecoderRef(cfg)
    ref.load_state_dict({k: v.float() for k, v in sd.items()})
    out = {}
    for res in (512, 1024):
        lat = res // 8
        z = (torch.randn(1, 4, lat, lat) * cfg.scaling_factor).half().cuda()
        with torch.no_grad():
            out[f"native_ms_{res}"] = ev_time(lambda: native.decode(z))
            r16 = ref.half().cuda()
            out[f"eager_fp16_ms_{res}"] = ev_time(lambda: r16.decode(z), iters=3)
            r32 = ref.float().cuda()
            out[f"eager_fp32_ms_{res}"] = ev_time(lambda: r32.decode(z.float()), iters=2)
        print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "vae_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
