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


class HashTokenPromptEncoder:
    """The two SDXL text encoders (CLIP ViT-L + OpenCLIP bigG, real architectures, random-init weights made on the GPU) on
    the native CLIP towers.  No BPE vocabulary exists offline, so words are hashed to token ids -- the arithmetic per prompt
    is exactly that of encode_prompt (ip_adapter/encoders.py::ClipPromptEncoder without the tokenizer files)."""

    def __init__(self, device):
        from imagharmony_b200.clip import ClipTextTower, ClipTowerConfig, tower_param_shapes
        cl = ClipTowerConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                             hidden_act="quick_gelu", projection_dim=None, vocab_size=49408, eos_token_id=49407)
        cg = ClipTowerConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                             hidden_act="gelu", projection_dim=1280, vocab_size=49408, eos_token_id=49407)
        self.towers = [ClipTextTower(c, random_state_dict(tower_param_shapes(c, "text"), 40 + i, device=device), device=device)
                       for i, c in enumerate((cl, cg))]

    @staticmethod
    def ids(prompts):
        import zlib
        out = torch.full((len(prompts), 77), 49407, dtype=torch.int64)
        for b, p in enumerate(prompts):
            toks = [1000 + zlib.crc32(w.encode()) % 40000 for w in p.replace(",", " ").split()][:75]
            out[b, 0] = 49406
            out[b, 1:1 + len(toks)] = torch.tensor(toks, dtype=torch.int64)
        return out

    def __call__(self, prompts):
        ids = self.ids(list(prompts))
        embeds, pooled = [], None
        for t in self.towers:
            o = t(ids)
            pooled = o.text_embeds if o.text_embeds is not None else o.pooler_output
            embeds.append(o.penultimate)
        return torch.cat(embeds, dim=-1), pooled


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

    # the same edit with the conditioning encoders inside the timed region: PIL image -> CLIPImageProcessor (host) -> native
    # ViT-bigG/14 -> image embeds; prompt / negative prompt / auxiliary text -> both native text towers (row f2)
    from PIL import Image
    from imagharmony_b200.clip import ClipTowerConfig, ClipVisionTower, tower_param_shapes
    vcfg = ClipTowerConfig(hidden_size=1664, intermediate_size=8192, num_hidden_layers=48, num_attention_heads=16,
                           hidden_act="gelu", projection_dim=1280, image_size=224, patch_size=14)
    vision = ClipVisionTower(vcfg, random_state_dict(tower_param_shapes(vcfg, "vision"), 31, device="cuda"), device="cuda")
    pipe.prompt_encoder = HashTokenPromptEncoder("cuda")
    ip2 = IPAdapterXL(pipe, vision, None, "cuda", num_tokens=4, target_blocks=["down_blocks.2.attentions.1"], inference=True,
                      number_class_crossattention=ha)
    ip2.image_proj_model = ip.image_proj_model
    procs = torch.nn.ModuleList(pipe.unet.attn_processors.values())       # set_ip_adapter installed fresh (zeroed) processors
    procs.load_state_dict({k: v.cuda() for k, v in random_state_dict(shapes_of(procs), 5).items()})
    pipe.unet.finalize()
    src = Image.fromarray((torch.rand(1024, 1024, 3, generator=torch.Generator("cpu").manual_seed(6)) * 255).byte().numpy())
    kw2 = dict(kw)
    kw2.pop("clip_image_embeds")
    kw2["pil_image"] = src
    ip2.generate(seed=[1], **kw2)
    ts = []
    for sd in (2, 3, 4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ip2.generate(seed=[sd], **kw2)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    out["latency_s_pil_with_native_encoders"] = sorted(ts)[1]
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "edit_latency.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
