"""StableDiffusionXLCustomPipeline with the reference's call surface (ip_adapter/custom_pipelines.py:16-394) on the
native UNet / DenoiseEngine.  The reference class derives from diffusers' StableDiffusionXLPipeline; diffusers is not
a dependency here, so the pieces the reference touches are provided directly:

    .unet (.config.{cross_attention_dim, block_out_channels, in_channels}, .attn_processors, .set_attn_processor)
    .scheduler, .set_scale(scale), .to(device), .enable_vae_tiling(), .encode_prompt(...), .__call__(...)

Denoise loop semantics follow custom_pipelines.py:188-363 (CFG batch [negative, positive], per-step IP-scale gating
by control_guidance_start/end, Euler update); the loop itself runs as replayed CUDA graphs (imagharmony_b200/denoise.py).
VAE decode / PIL post-processing (:365-386, scope row f1) runs on the native decoder `imagharmony_b200.vae` when the
pipeline owns one (`from_random`, or `from_pretrained` with a `vae/` folder); a `vae_decode` callable overrides it and
`output_type="latent"` skips it.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch

from imagharmony_b200._lib import IHError
from imagharmony_b200.config import SDXL_BASE, UNetConfig
from imagharmony_b200.denoise import DenoiseEngine
from imagharmony_b200.scheduler import EulerDiscreteScheduler
from imagharmony_b200.unet import UNet2DConditionModel

from .utils import is_torch2_available

if is_torch2_available():
    from .attention_processor import IPAttnProcessor2_0 as IPAttnProcessor
else:  # pragma: no cover
    from .attention_processor import IPAttnProcessor


@dataclass
class StableDiffusionXLPipelineOutput:
    images: Any


class SyntheticPromptEncoder:
    """Stand-in for the two CLIP text encoders when no weights are on disk (this environment has no network):
    deterministic N(0,1) embeddings seeded by a hash of the prompt string.  Shapes/dtypes match encode_prompt [3P]."""

    def __init__(self, cfg: UNetConfig):
        self.cfg = cfg

    def __call__(self, prompts: List[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        hs, pooled = [], []
        for p in prompts:
            seed = int.from_bytes(hashlib.sha256(p.encode()).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
            g = torch.Generator("cpu").manual_seed(seed)
            hs.append(torch.randn(77, self.cfg.cross_attention_dim, generator=g))
            pooled.append(torch.randn(self.cfg.pooled_embed_dim, generator=g))
        return torch.stack(hs).half(), torch.stack(pooled).half()


class StableDiffusionXLCustomPipeline:
    vae_scale_factor = 8
    force_zeros_for_empty_prompt = True     # [3P] SDXL-base pipeline config

    def __init__(self, unet: UNet2DConditionModel, prompt_encoder: Optional[Callable] = None,
                 vae_decode: Optional[Callable] = None, scheduler: Optional[EulerDiscreteScheduler] = None,
                 vae=None):
        self.unet = unet
        self.vae = vae                      # imagharmony_b200.vae.AutoencoderKLDecoder or None
        self.scheduler = scheduler or EulerDiscreteScheduler()
        self.prompt_encoder = prompt_encoder or SyntheticPromptEncoder(unet.config)
        self.vae_decode = vae_decode
        self.default_sample_size = unet.config.sample_size
        self._engine: Optional[DenoiseEngine] = None
        self.device = unet.conv_in.weight.device

    # ---- construction ------------------------------------------------------------------------------------------
    @classmethod
    def from_random(cls, cfg: UNetConfig = SDXL_BASE, seed: int = 0, device="cuda", vae_cfg=None):
        """Random-init weights of the given architecture (the benchmark / test configuration)."""
        from imagharmony_b200.weights import random_state_dict, shapes_of
        with torch.device("meta"):
            shapes = shapes_of(UNet2DConditionModel(cfg))
        gen_dev = device if str(device).startswith("cuda") else "cpu"
        unet = UNet2DConditionModel.from_state_dict(cfg, random_state_dict(shapes, seed, device=gen_dev), device=device)
        vae = None
        if vae_cfg is not None:
            from imagharmony_b200.vae import AutoencoderKLDecoder
            with torch.device("meta"):
                vshapes = shapes_of(AutoencoderKLDecoder(vae_cfg))
            vae = AutoencoderKLDecoder.from_state_dict(vae_cfg, random_state_dict(vshapes, seed + 17, device=gen_dev),
                                                       device=device)
        return cls(unet, vae=vae)

    @classmethod
    def from_pretrained(cls, path: str, torch_dtype=torch.float16, add_watermarker: bool = False, device="cuda",
                        cfg: UNetConfig = SDXL_BASE, vae_cfg=None, **kwargs):
        """test.py:68-72.  Loads `<path>/unet/diffusion_pytorch_model.safetensors` (diffusers key names are the native
        UNet's key names, so no key map is needed)."""
        from safetensors.torch import load_file
        f = os.path.join(path, "unet", "diffusion_pytorch_model.safetensors")
        if not os.path.exists(f):
            f = os.path.join(path, "unet", "diffusion_pytorch_model.fp16.safetensors")
        if not os.path.exists(f):
            raise FileNotFoundError(f"no SDXL UNet weights under {path}/unet (safetensors)")
        unet = UNet2DConditionModel.from_state_dict(cfg, load_file(f), device=device)
        vae = None
        for name in ("diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.safetensors"):
            vf = os.path.join(path, "vae", name)
            if os.path.exists(vf):
                import dataclasses
                import json
                from imagharmony_b200.config import SDXL_VAE
                from imagharmony_b200.vae import AutoencoderKLDecoder
                vcfg = vae_cfg or SDXL_VAE
                cj = os.path.join(path, "vae", "config.json")
                if vae_cfg is None and os.path.exists(cj):
                    # `force_upcast` decides between the scaled residual stream (the reference's fp32 upcast,
                    # custom_pipelines.py:366-371) and plain fp16 (fp16-fix checkpoints set it to false)
                    meta = json.load(open(cj))
                    vcfg = dataclasses.replace(vcfg, force_upcast=bool(meta.get("force_upcast", True)),
                                               scaling_factor=float(meta.get("scaling_factor", vcfg.scaling_factor)))
                vae = AutoencoderKLDecoder.from_state_dict(vcfg, load_file(vf), device=device)  # decoder keys only
                break
        enc = None
        from .encoders import ClipPromptEncoder
        if ClipPromptEncoder.available(path):
            enc = ClipPromptEncoder.from_pretrained(path, device=device, dtype=torch_dtype)
        return cls(unet, prompt_encoder=enc, vae=vae)

    def to(self, device=None, *args, **kwargs):
        """Weights live on the device the UNet was built on (from_random / from_pretrained take `device=`); moving 5 GB
        of kernel-layout weights behind the caller's back is not offered, a mismatching request raises."""
        if device is not None and not isinstance(device, torch.dtype):
            want = torch.device(device)
            have = torch.device(self.device)
            if want.type != have.type or (want.index is not None and have.index is not None and want.index != have.index):
                raise IHError(f"pipeline was built on {have}; build it with device={want!s} instead of .to({want!s})")
        return self

    def enable_vae_tiling(self):  # test.py:73
        if self.vae is not None:
            self.vae.use_tiling = True

    # ---- reference surface -------------------------------------------------------------------------------------
    def set_scale(self, scale):
        for attn_processor in self.unet.attn_processors.values():           # custom_pipelines.py:17-20
            if isinstance(attn_processor, IPAttnProcessor):
                attn_processor.scale = scale

    @property
    def engine(self) -> DenoiseEngine:
        if self._engine is None or self._engine.scheduler is not self.scheduler:
            # the loop integrates the SAME scheduler object whose init_noise_sigma scaled the initial latents
            self._engine = DenoiseEngine(self.unet, scheduler=self.scheduler)
        return self._engine

    @torch.no_grad()
    def encode_prompt(self, prompt, prompt_2=None, device=None, num_images_per_prompt: int = 1,
                      do_classifier_free_guidance: bool = True, negative_prompt=None, negative_prompt_2=None,
                      prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                      negative_pooled_prompt_embeds=None, lora_scale=None, **kwargs):
        """-> (prompt_embeds [n,77,2048], negative_prompt_embeds, pooled [n,1280], negative_pooled)  ([3P] A.4)."""
        if prompt_embeds is None:
            prompts = [prompt] if isinstance(prompt, str) else list(prompt)
            prompt_embeds, pooled_prompt_embeds = self.prompt_encoder(prompts)
            prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
            pooled_prompt_embeds = pooled_prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
        if do_classifier_free_guidance and negative_prompt_embeds is None and negative_prompt is None \
                and self.force_zeros_for_empty_prompt:
            # [3P] SDXL-base model_index.json: force_zeros_for_empty_prompt = true -> no negative prompt means ZERO embeds
            negative_prompt_embeds = torch.zeros_like(prompt_embeds)
            negative_pooled_prompt_embeds = torch.zeros_like(pooled_prompt_embeds)
        elif do_classifier_free_guidance and negative_prompt_embeds is None:
            negs = negative_prompt if negative_prompt is not None else ""
            negs = [negs] * (prompt_embeds.shape[0] // num_images_per_prompt) if isinstance(negs, str) else list(negs)
            negative_prompt_embeds, negative_pooled_prompt_embeds = self.prompt_encoder(negs)
            negative_prompt_embeds = negative_prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
            negative_pooled_prompt_embeds = negative_pooled_prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
        return prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        """[3P] diffusers prepare_latents: randn * init_noise_sigma; a list of generators draws one sample each."""
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            if isinstance(generator, list):
                if len(generator) != batch_size:
                    raise ValueError(f"got {len(generator)} generators for a batch of {batch_size}")
                parts = [torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=torch.float32) for g in generator]
                latents = torch.cat([p.to("cpu") for p in parts], dim=0)
            else:
                gdev = generator.device if generator is not None else "cpu"
                latents = torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32).to("cpu")
        latents = latents.to(torch.float32) * self.scheduler.init_noise_sigma
        return latents.to(dtype)

    @torch.no_grad()
    def __call__(self, prompt: Optional[Union[str, List[str]]] = None, prompt_2=None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, denoising_end: Optional[float] = None,
                 guidance_scale: float = 5.0, negative_prompt=None, negative_prompt_2=None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None,
                 latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None,
                 pooled_prompt_embeds: Optional[torch.Tensor] = None,
                 negative_pooled_prompt_embeds: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback=None, callback_steps: int = 1,
                 cross_attention_kwargs: Optional[Dict[str, Any]] = None, guidance_rescale: float = 0.0,
                 original_size: Optional[Tuple[int, int]] = None, crops_coords_top_left: Tuple[int, int] = (0, 0),
                 target_size: Optional[Tuple[int, int]] = None, negative_original_size=None,
                 negative_crops_coords_top_left=(0, 0), negative_target_size=None,
                 control_guidance_start: float = 0.0, control_guidance_end: float = 1.0, **ignored):
        """Parameter list of custom_pipelines.py:23-56.  Unknown keyword arguments are accepted and ignored, because
        IPAdapterXL.generate forwards a stray `number_class_crossattention=` into this call (test.py:38, demo.py:124)."""
        height = height or self.default_sample_size * self.vae_scale_factor           # :189-190
        width = width or self.default_sample_size * self.vae_scale_factor
        original_size = original_size or (height, width)                              # :192-193
        target_size = target_size or (height, width)
        if callback is not None and (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps}")   # [3P] check_inputs
        do_classifier_free_guidance = guidance_scale > 1.0                            # :223
        (prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds) = \
            self.encode_prompt(prompt, prompt_2, None, num_images_per_prompt, do_classifier_free_guidance,
                               negative_prompt, negative_prompt_2, prompt_embeds, negative_prompt_embeds,
                               pooled_prompt_embeds, negative_pooled_prompt_embeds)  # :229-247
        batch = prompt_embeds.shape[0]
        self.scheduler.set_timesteps(num_inference_steps)                             # :250
        lat = self.prepare_latents(batch, self.unet.config.in_channels, height, width, torch.float16, self.device,
                                   generator, latents)                               # :255-265

        def time_ids(size, crop, target):                                             # :277-284 _get_add_time_ids [3P]
            return torch.tensor([list(size) + list(crop) + list(target)], dtype=torch.float32).repeat(batch, 1)
        add_time_ids = time_ids(original_size, crops_coords_top_left, target_size)
        negative_add_time_ids = None
        if negative_original_size is not None and negative_target_size is not None:   # :285-294
            negative_add_time_ids = time_ids(negative_original_size, negative_crops_coords_top_left, negative_target_size)
        loop_steps = num_inference_steps
        if denoising_end is not None and isinstance(denoising_end, float) and 0 < denoising_end < 1:   # :307-316
            n_train = self.scheduler.num_train_timesteps
            cutoff = int(round(n_train - denoising_end * n_train))
            loop_steps = int(sum(1 for t in self.scheduler.timesteps if t >= cutoff))
        conditioning_scale = 1.0
        for attn_processor in self.unet.attn_processors.values():                     # :319-322
            if isinstance(attn_processor, IPAttnProcessor):
                conditioning_scale = attn_processor.scale
                break
        if loop_steps > 0:
            out = self.engine.run(lat, prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds,
                                  negative_pooled_prompt_embeds, add_time_ids, num_inference_steps,
                                  guidance_scale=guidance_scale, ip_scale=conditioning_scale,
                                  control_guidance_start=control_guidance_start,
                                  control_guidance_end=control_guidance_end, guidance_rescale=guidance_rescale,
                                  num_loop_steps=loop_steps, callback=callback, callback_steps=callback_steps,
                                  negative_time_ids=negative_add_time_ids)
        else:
            out = lat.to(self.device)
        self.set_scale(conditioning_scale)
        if output_type == "latent":
            image = out
        else:
            if self.vae_decode is not None:
                image = self.vae_decode(out / 0.13025, output_type)                   # :373 (scaling_factor [3P])
            elif self.vae is not None:
                from imagharmony_b200.vae import postprocess
                image = postprocess(self.vae.decode(out), output_type)                # :373 + :383 (/ scaling_factor folded)
            else:
                raise IHError("this pipeline has no VAE decoder: call with output_type='latent', build it with a `vae` "
                              "(from_random(vae_cfg=...), from_pretrained with a vae/ folder) or pass a vae_decode callable")
        if not return_dict:
            return (image,)
        return StableDiffusionXLPipelineOutput(images=image)
