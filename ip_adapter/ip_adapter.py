"""IPAdapter / IPAdapterXL / IPAdapterPlusXL with the reference's constructor and method surface
(reference ip_adapter/ip_adapter.py:69-340, 389-478) driving the native pipeline.

What is preserved: argument names/defaults, attribute names (.pipe, .image_encoder, .clip_image_processor,
.image_proj_model, .number_class_crossattention, .device, .num_tokens, .generator), set_ip_adapter's layer selection
(IP branch only where the processor name contains 'down_blocks.2.attentions.1', :117; `target_blocks` is accepted and
ignored exactly like the reference, :75), the 3-key checkpoint layout of load_ip_adapter (:149-154), get_image_embeds,
set_scale and generate().

Positions taken on reference quirks: `extra_text=None` skips the HarmonyAttention residual
instead of raising NameError (C.5); the .safetensors branch is implemented correctly instead of KeyError-ing (C.6);
the forward-time prints of HarmonyAttention are dropped (C.10); one copy of the auxiliary text feeds HA (C.11).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from PIL import Image

from imagharmony_b200._lib import IHError
from imagharmony_b200.adapter import HarmonyAttention, ImageProjModel, Resampler  # noqa: F401

from .utils import get_generator, is_torch2_available

if is_torch2_available():
    from .attention_processor import AttnProcessor2_0 as AttnProcessor
    from .attention_processor import IPAttnProcessor2_0 as IPAttnProcessor
else:  # pragma: no cover
    from .attention_processor import AttnProcessor, IPAttnProcessor


class IPAdapter:
    def __init__(self, sd_pipe, image_encoder_path, ip_ckpt, device, num_tokens=4, target_blocks=None,
                 number_class_crossattention=None):
        self.device = device
        self.image_encoder_path = image_encoder_path
        self.ip_ckpt = ip_ckpt
        self.num_tokens = num_tokens
        self.pipe = sd_pipe.to(self.device)
        self.set_ip_adapter()

        # image encoder (CLIP ViT-bigG/14 with projection for SDXL): a [3P] transformers model, "next" row f2
        self.image_encoder = None
        self.clip_image_processor = None
        if image_encoder_path is not None and os.path.isdir(str(image_encoder_path)):
            from transformers import CLIPImageProcessor, CLIPVisionModelWithProjection
            self.image_encoder = CLIPVisionModelWithProjection.from_pretrained(image_encoder_path).to(
                self.device, dtype=torch.float16)
            self.clip_image_processor = CLIPImageProcessor()
        self.number_class_crossattention = None
        if number_class_crossattention is not None:
            self.number_class_crossattention = number_class_crossattention.to(self.device, dtype=torch.float16)
        self.image_proj_model = self.init_proj()
        self.generator = None
        self.load_ip_adapter()

    # ---- reference :91-97 -------------------------------------------------------------------------------------------
    def _clip_dim(self) -> int:
        if self.image_encoder is not None:
            return self.image_encoder.config.projection_dim
        if self.number_class_crossattention is not None:
            return self.number_class_crossattention.image_hidden_size
        return self.pipe.unet.config.pooled_embed_dim

    def init_proj(self):
        image_proj_model = ImageProjModel(
            cross_attention_dim=self.pipe.unet.config.cross_attention_dim,
            clip_embeddings_dim=self._clip_dim(),
            clip_extra_context_tokens=self.num_tokens,
        ).to(self.device, dtype=torch.float16)
        return image_proj_model.requires_grad_(False)

    # ---- reference :99-133 ------------------------------------------------------------------------------------------
    def set_ip_adapter(self):
        unet = self.pipe.unet
        attn_procs = {}
        for name in unet.attn_processors.keys():
            cross_attention_dim = None if name.endswith("attn1.processor") else unet.config.cross_attention_dim
            if name.startswith("mid_block"):
                hidden_size = unet.config.block_out_channels[-1]
            elif name.startswith("up_blocks"):
                block_id = int(name[len("up_blocks.")])
                hidden_size = list(reversed(unet.config.block_out_channels))[block_id]
            elif name.startswith("down_blocks"):
                block_id = int(name[len("down_blocks.")])
                hidden_size = unet.config.block_out_channels[block_id]
            if cross_attention_dim is None:
                attn_procs[name] = AttnProcessor()
            else:
                skip = "down_blocks.2.attentions.1" not in name                      # :117-123
                with torch.device("meta"):
                    proc = IPAttnProcessor(hidden_size=hidden_size, cross_attention_dim=cross_attention_dim,
                                           num_tokens=self.num_tokens, skip=skip)
                proc = proc.to_empty(device=self.device).to(torch.float16).requires_grad_(False)
                for p in proc.parameters():
                    p.zero_()
                attn_procs[name] = proc
        unet.set_attn_processor(attn_procs)
        if hasattr(self.pipe, "controlnet"):
            raise IHError("ControlNet pipelines are outside the SDXL IP-adapter hot path")

    # ---- reference :135-154 -----------------------------------------------------------------------------------------
    def load_ip_adapter(self):
        if self.ip_ckpt is None:
            return   # random / zero-initialised adapter (benchmarks and tests: there are no checkpoints offline)
        if os.path.splitext(self.ip_ckpt)[-1] == ".safetensors":
            from safetensors import safe_open
            state_dict = {"image_proj": {}, "ip_adapter": {}, "composed_adapter": {}}
            with safe_open(self.ip_ckpt, framework="pt", device="cpu") as f:
                for key in f.keys():
                    for prefix, dst in (("image_proj.", "image_proj"), ("image_proj_model.", "image_proj"),
                                        ("ip_adapter.", "ip_adapter"), ("adapter_modules.", "ip_adapter"),
                                        ("composed_adapter.", "composed_adapter"),
                                        ("composed_modules.", "composed_adapter")):
                        if key.startswith(prefix):
                            state_dict[dst][key[len(prefix):]] = f.get_tensor(key)
                            break
        else:
            state_dict = torch.load(self.ip_ckpt, map_location="cpu")
        self.image_proj_model.load_state_dict(state_dict["image_proj"])
        if self.number_class_crossattention is not None and state_dict.get("composed_adapter"):
            self.number_class_crossattention.load_state_dict(state_dict["composed_adapter"])
        ip_layers = torch.nn.ModuleList(self.pipe.unet.attn_processors.values())
        ip_layers.load_state_dict(state_dict["ip_adapter"])       # strict, index-keyed: all 70 attn2 own tensors (:153)
        self.pipe.unet.finalize()

    # ---- reference :158-182 -----------------------------------------------------------------------------------------
    @torch.inference_mode()
    def get_image_embeds(self, pil_image=None, clip_image_embeds=None, extra_prompt_embeds=None):
        if pil_image is not None:
            if self.image_encoder is None:
                raise IHError("no CLIP image encoder loaded (image_encoder_path): pass clip_image_embeds=")
            if isinstance(pil_image, Image.Image):
                pil_image = [pil_image]
            clip_image = self.clip_image_processor(images=pil_image, return_tensors="pt").pixel_values
            clip_image_embeds = self.image_encoder(clip_image.to(self.device, dtype=torch.float16)).image_embeds
        else:
            clip_image_embeds = clip_image_embeds.to(self.device, dtype=torch.float16)
        clip_image_embeds = clip_image_embeds.contiguous()
        if extra_prompt_embeds is not None and self.number_class_crossattention is not None:      # :169-173
            extra_prompt_embeds = extra_prompt_embeds.to(self.device, torch.float16)
            clip_image_embeds = self.number_class_crossattention(extra_prompt_embeds, clip_image_embeds,
                                                                 add_to=clip_image_embeds)
        image_prompt_embeds = self.image_proj_model(clip_image_embeds)                            # :175
        uncond_image_prompt_embeds = self.image_proj_model(torch.zeros_like(clip_image_embeds))  # :176
        return image_prompt_embeds, uncond_image_prompt_embeds

    def set_scale(self, scale):
        for attn_processor in self.pipe.unet.attn_processors.values():
            if isinstance(attn_processor, IPAttnProcessor):
                attn_processor.scale = scale


class IPAdapterXL(IPAdapter):
    """SDXL"""

    def __init__(self, sd_pipe, image_encoder_path, ip_ckpt, device, num_tokens=4, target_blocks=None,
                 inference=False, number_class_crossattention=None):
        self.inference = inference
        super().__init__(sd_pipe, image_encoder_path, ip_ckpt, device, num_tokens=num_tokens,
                         target_blocks=target_blocks, number_class_crossattention=number_class_crossattention)

    def generate(self, pil_image=None, prompt=None, negative_prompt=None, extra_text=None, scale=1.0, num_samples=4,
                 seed=None, num_inference_steps=30, clip_image_embeds=None, **kwargs):
        """reference :257-340.  `clip_image_embeds=` is an extension for environments without the CLIP vision weights."""
        self.set_scale(scale)
        if pil_image is not None:
            num_prompts = 1 if isinstance(pil_image, Image.Image) else len(pil_image)
        else:
            num_prompts = clip_image_embeds.shape[0]
        if prompt is None:
            prompt = "best quality, high quality"
        if negative_prompt is None:
            negative_prompt = "monochrome, lowres, bad anatomy, worst quality, low quality"
        if not isinstance(prompt, List):
            prompt = [prompt] * num_prompts
        if not isinstance(negative_prompt, List):
            negative_prompt = [negative_prompt] * num_prompts

        extra_prompt_embeds = None
        if extra_text is not None:                                                                # :285-297
            extra_prompt_embeds, _, _, _ = self.pipe.encode_prompt(
                extra_text, num_images_per_prompt=num_samples, do_classifier_free_guidance=True,
                negative_prompt=negative_prompt)
        image_prompt_embeds, uncond_image_prompt_embeds = self.get_image_embeds(
            pil_image=pil_image, clip_image_embeds=clip_image_embeds, extra_prompt_embeds=extra_prompt_embeds)  # :300

        bs_embed, seq_len, _ = image_prompt_embeds.shape                                          # :302-306
        image_prompt_embeds = image_prompt_embeds.repeat(1, num_samples, 1).view(bs_embed * num_samples, seq_len, -1)
        uncond_image_prompt_embeds = uncond_image_prompt_embeds.repeat(1, num_samples, 1).view(
            bs_embed * num_samples, seq_len, -1)

        (prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds,
         negative_pooled_prompt_embeds) = self.pipe.encode_prompt(
            prompt, num_images_per_prompt=num_samples, do_classifier_free_guidance=True,
            negative_prompt=negative_prompt)                                                      # :308-319
        dev = image_prompt_embeds.device
        prompt_embeds = torch.cat([prompt_embeds.to(dev), image_prompt_embeds], dim=1)            # :321
        negative_prompt_embeds = torch.cat([negative_prompt_embeds.to(dev), uncond_image_prompt_embeds], dim=1)  # :322

        gen_device = "cpu" if isinstance(seed, list) else self.device
        self.generator = get_generator(seed, gen_device)                                          # :328
        images = self.pipe(prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds,
                           pooled_prompt_embeds=pooled_prompt_embeds,
                           negative_pooled_prompt_embeds=negative_pooled_prompt_embeds,
                           num_inference_steps=num_inference_steps, generator=self.generator, **kwargs).images
        return images


class IPAdapterPlusXL(IPAdapter):
    """SDXL with the Resampler projector (reference :389-478): 16 image tokens from the CLIP penultimate hidden
    states."""

    def init_proj(self):
        emb = self.image_encoder.config.hidden_size if self.image_encoder is not None else 1664
        image_proj_model = Resampler(dim=1280, depth=4, dim_head=64, heads=20, num_queries=self.num_tokens,
                                     embedding_dim=emb, output_dim=self.pipe.unet.config.cross_attention_dim,
                                     ff_mult=4).to(self.device, dtype=torch.float16)
        return image_proj_model.requires_grad_(False)

    @torch.inference_mode()
    def get_image_embeds(self, pil_image=None, clip_hidden_states=None, uncond_clip_hidden_states=None):
        if pil_image is not None:
            if self.image_encoder is None:
                raise IHError("no CLIP image encoder loaded: pass clip_hidden_states= / uncond_clip_hidden_states=")
            if isinstance(pil_image, Image.Image):
                pil_image = [pil_image]
            clip_image = self.clip_image_processor(images=pil_image, return_tensors="pt").pixel_values
            clip_image = clip_image.to(self.device, dtype=torch.float16)
            clip_hidden_states = self.image_encoder(clip_image, output_hidden_states=True).hidden_states[-2]
            uncond_clip_hidden_states = self.image_encoder(torch.zeros_like(clip_image),
                                                           output_hidden_states=True).hidden_states[-2]
        image_prompt_embeds = self.image_proj_model(clip_hidden_states.to(self.device, torch.float16))
        uncond_image_prompt_embeds = self.image_proj_model(uncond_clip_hidden_states.to(self.device, torch.float16))
        return image_prompt_embeds, uncond_image_prompt_embeds

    def generate(self, pil_image=None, prompt=None, negative_prompt=None, scale=1.0, num_samples=4, seed=None,
                 num_inference_steps=30, clip_hidden_states=None, uncond_clip_hidden_states=None, **kwargs):
        self.set_scale(scale)
        num_prompts = 1 if (pil_image is None or isinstance(pil_image, Image.Image)) else len(pil_image)
        if prompt is None:
            prompt = "best quality, high quality"
        if negative_prompt is None:
            negative_prompt = "monochrome, lowres, bad anatomy, worst quality, low quality"
        if not isinstance(prompt, List):
            prompt = [prompt] * num_prompts
        if not isinstance(negative_prompt, List):
            negative_prompt = [negative_prompt] * num_prompts
        image_prompt_embeds, uncond_image_prompt_embeds = self.get_image_embeds(
            pil_image, clip_hidden_states, uncond_clip_hidden_states)
        bs_embed, seq_len, _ = image_prompt_embeds.shape
        image_prompt_embeds = image_prompt_embeds.repeat(1, num_samples, 1).view(bs_embed * num_samples, seq_len, -1)
        uncond_image_prompt_embeds = uncond_image_prompt_embeds.repeat(1, num_samples, 1).view(
            bs_embed * num_samples, seq_len, -1)
        (prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds,
         negative_pooled_prompt_embeds) = self.pipe.encode_prompt(
            prompt, num_images_per_prompt=num_samples, do_classifier_free_guidance=True,
            negative_prompt=negative_prompt)
        dev = image_prompt_embeds.device
        prompt_embeds = torch.cat([prompt_embeds.to(dev), image_prompt_embeds], dim=1)
        negative_prompt_embeds = torch.cat([negative_prompt_embeds.to(dev), uncond_image_prompt_embeds], dim=1)
        generator = get_generator(seed, "cpu" if isinstance(seed, list) else self.device)
        return self.pipe(prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds,
                         pooled_prompt_embeds=pooled_prompt_embeds,
                         negative_pooled_prompt_embeds=negative_pooled_prompt_embeds,
                         num_inference_steps=num_inference_steps, generator=generator, **kwargs).images
