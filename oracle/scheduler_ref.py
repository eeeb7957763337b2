"""ORACLE (test infrastructure only). Restatement of the scheduler + denoise loop the reference drives.

EulerDiscreteScheduler lives in diffusers==0.30.0 ([3P], requirements.txt:25; call sites custom_pipelines.py:250,
268,334,357).  SDXL-base scheduler_config: scaled_linear betas 0.00085 -> 0.012, 1000 train steps, epsilon
prediction, timestep_spacing="leading", steps_offset=1.  Known-answer values: SURVEY.md appendix A.3.
The loop body follows custom_pipelines.py:325-363.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch


def euler_tables(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, steps_offset: int = 1):
    """-> (timesteps float32 [T], sigmas float32 [T+1] with trailing 0, init_noise_sigma float)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    sig_all = (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).numpy()
    ratio = num_train_timesteps // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32) + steps_offset
    sig = np.interp(ts, np.arange(0, len(sig_all)), sig_all)
    sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
    init_noise_sigma = float((sigmas.max() ** 2 + 1) ** 0.5)          # "leading" spacing
    return ts.astype(np.float32), sigmas, init_noise_sigma


def prepare_latents(n: int, channels: int, h: int, w: int, seeds: Sequence[int], init_noise_sigma: float,
                    dtype=torch.float16) -> torch.Tensor:
    """One CPU generator per image (ip_adapter/utils.py:83-93 list-of-seeds semantics; diffusers randn_tensor draws
    one [1,C,h,w] tensor per generator) so a candidate noise is identical whichever rank / batch slot holds it."""
    assert len(seeds) == n
    parts = [torch.randn((1, channels, h, w), generator=torch.Generator("cpu").manual_seed(int(s)), dtype=torch.float32)
             for s in seeds]
    return (torch.cat(parts, 0) * init_noise_sigma).to(dtype)


def rescale_noise_cfg(noise_cfg: torch.Tensor, noise_pred_text: torch.Tensor, guidance_rescale: float) -> torch.Tensor:
    """[3P] diffusers==0.30.0 pipeline_stable_diffusion_xl.rescale_noise_cfg (called at custom_pipelines.py:352-354;
    Lin et al. 2023, section 3.4): match the per-image std of the guided prediction to the text branch's, blend."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


def denoising_end_steps(num_inference_steps: int, denoising_end: Optional[float], num_train_timesteps: int = 1000) -> int:
    """custom_pipelines.py:307-316: number of leading timesteps kept when `denoising_end` is a float in (0, 1)."""
    if denoising_end is None or not isinstance(denoising_end, float) or not 0 < denoising_end < 1:
        return num_inference_steps
    timesteps, _, _ = euler_tables(num_inference_steps, num_train_timesteps)
    cutoff = int(round(num_train_timesteps - denoising_end * num_train_timesteps))
    return int(sum(1 for t in timesteps if t >= cutoff))


@torch.no_grad()
def denoise_loop(unet_fn: Callable, latents: torch.Tensor, prompt_embeds: torch.Tensor, neg_prompt_embeds: torch.Tensor,
                 pooled: torch.Tensor, neg_pooled: torch.Tensor, time_ids: torch.Tensor, num_inference_steps: int,
                 guidance_scale: float = 5.0, set_scale: Optional[Callable[[float], None]] = None,
                 conditioning_scale: float = 1.0, control_guidance_start: float = 0.0,
                 control_guidance_end: float = 1.0, trace: Optional[list] = None, guidance_rescale: float = 0.0,
                 denoising_end: Optional[float] = None, callback: Optional[Callable] = None,
                 callback_steps: int = 1) -> torch.Tensor:
    """custom_pipelines.py:296-363.  `unet_fn(sample, t, ehs, text_embeds, time_ids) -> noise_pred`.
    Arithmetic dtype follows the tensors' dtype exactly like the reference (fp16 tensors -> fp16 rounding points;
    the Euler step itself is fp32 inside and cast back, [3P] EulerDiscreteScheduler.step).
    guidance_scale <= 1 disables classifier-free guidance (:223): positive branch only."""
    timesteps, sigmas, _ = euler_tables(num_inference_steps)
    cfg = guidance_scale > 1.0                                                    # :223
    if cfg:
        ehs = torch.cat([neg_prompt_embeds, prompt_embeds], dim=0)                # :296
        text_embeds = torch.cat([neg_pooled, pooled], dim=0)                      # :297
        tids = torch.cat([time_ids, time_ids], dim=0)                             # :298, :302
    else:
        ehs, text_embeds, tids = prompt_embeds, pooled, time_ids
    T = denoising_end_steps(num_inference_steps, denoising_end)                   # :307-316 (timesteps[:T])
    for i in range(T):
        if set_scale is not None:                                                 # :326-329
            off = (i / T < control_guidance_start) or ((i + 1) / T > control_guidance_end)
            set_scale(0.0 if off else conditioning_scale)
        sigma, sigma_next = float(sigmas[i]), float(sigmas[i + 1])
        x2 = torch.cat([latents] * 2) if cfg else latents                         # :332
        x2 = (x2 / ((sigma ** 2 + 1) ** 0.5)).to(latents.dtype)                   # :334 scale_model_input
        noise = unet_fn(x2, float(timesteps[i]), ehs, text_embeds, tids)          # :338-345
        if cfg:
            u, c = noise.chunk(2)                                                 # :349
            eps = u + guidance_scale * (c - u)                                    # :350
            if guidance_rescale > 0.0:                                            # :352-354
                eps = rescale_noise_cfg(eps, c, guidance_rescale)
        else:
            eps = noise
        x = latents.float()                                                       # Euler step (fp32 inside)
        # [3P] `sigma_hat * model_output`: 0-dim fp32 sigma times an fp16 tensor stays fp16 under torch type promotion
        x0 = x - (sigma * eps.float()).to(eps.dtype).float()
        deriv = (x - x0) / sigma
        latents = (x + deriv * (sigma_next - sigma)).to(eps.dtype)                # :357
        if trace is not None:
            trace.append(latents.clone())
        if callback is not None and i % callback_steps == 0:                      # :359-363
            callback(i, float(timesteps[i]), latents)
    return latents
