"""End-to-end edit latency through the reference's call surface (north_star: "1024^2 50-step edit latency"):
IPAdapterXL.generate(...) = HarmonyAttention -> ImageProjModel -> prompt embeds -> 50-step CUDA-graph denoise loop ->
native VAE decode -> PIL, random-init SDXL-base UNet / SDXL VAE / adapter weights, synthetic prompt + CLIP image
embeddings (no encoder weights offline).  Wall clock with a device synchronize on both sides.
Writes gpurun_out/edit_latency.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagharmony_b200.config import HARMONY_DEFAULT as h, SDXL_BASE, SDXL_VAE  # noqa: E402
from imagharmony_b200.weights import random_state_dict, shapes_of  # noqa: E402
from ip_adapter import IPAdapterXL  # noqa: E402
from ip_adapter.custom_pipelines import StableDiffusionXLCustomPipeline  # noqa: E402
from train import HarmonyAttention  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    pipe = StableDiffusionXLCustomPipeline.from_random(SDXL_BASE, seed=0, device="cuda", vae_cfg=SDXL_VAE)
    ha = HarmonyAttention(image_hidden_size=h.image_hidden_size, text_context_dim=h.text_context_dim, inter_dim=h.inter_dim,
                          cross_heads=h.cross_heads, reshape_blocks=h.reshape_blocks, cross_value_dim=h.cross_value_dim,
                          scale=1.0, fusion_method="cross_attention")
    ip = IPAdapterXL(pipe, None, None, "cuda", num_tokens=4, target_blocks=["down_blocks.2.attentions.1"], inference=True,
                     number_class_crossattention=ha)
    ip.image_proj_model.load_state_dict({k: v.cuda() for k, v in random_state_dict(shapes_of(ip.image_proj_model), 3).items()})
    ip.number_class_crossattention.load_state_dict({k: v.cuda() for k, v in random_state_dict(shapes_of(ha), 4).items()})
    procs = torch.nn.ModuleList(pipe.unet.attn_processors.values())
    procs.load_state_dict({k: v.cuda() for k, v in random_state_dict(shapes_of(procs), 5).items()})
    pipe.unet.finalize()
    img = torch.randn(1, h.image_hidden_size, generator=torch.Generator("cpu").manual_seed(5)).half()
    kw = dict(pil_image=None, clip_image_embeds=img, prompt="lions", negative_prompt="blurry", scale=1.0, guidance_scale=5.0,
              num_samples=1, num_inference_steps=steps, extra_text="eight sheep", height=1024, width=1024)
    out = {"steps": steps}
    ip.generate(seed=[1], **kw)                      # warm-up: graph capture, workspaces, allocator
    for name, extra in (("latency_s_pil", {}), ("latency_s_latent", {"output_type": "latent"})):
        ts = []
        for s in (2, 3, 4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = ip.generate(seed=[s], **kw, **extra)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        out[name] = sorted(ts)[1]
    out["image_size"] = list(res.shape[-2:]) if hasattr(res, "shape") else None
    pil = ip.generate(seed=[9], **kw)
    out["pil"] = [pil[0].size, pil[0].mode]
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "edit_latency.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
