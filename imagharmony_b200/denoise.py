"""The denoise loop of custom_pipelines.py:325-363 as a device-resident loop.

One step = UNet forward on the scaled CFG batch + the fused CFG/Euler kernel; everything a step needs (timestep,
sigma pair, step index) is read from device memory, so the step is captured ONCE as a CUDA graph and replayed T times
with no host work in between.  Step-invariant work (cross-attention K/V of all 70 attn2 layers, the text_time
embedding) runs once per call in `prepare`.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops
from ._lib import IHError
from .scheduler import EulerDiscreteScheduler
from .unet import UNet2DConditionModel


class DenoiseEngine:
    def __init__(self, unet: UNet2DConditionModel, use_cuda_graph: bool = True):
        self.unet = unet
        self.device = unet.conv_in.weight.device
        self.scheduler = EulerDiscreteScheduler()
        self.use_cuda_graph = use_cuda_graph and self.device.type == "cuda"
        self._graphs: Dict[Tuple, torch.cuda.CUDAGraph] = {}
        self._static: Dict[Tuple, dict] = {}
        self._tables: Dict[int, Tuple[torch.Tensor, torch.Tensor, float]] = {}
        self.last_launches_per_step = 0

    # ------------------------------------------------------------------------------------------------------------
    def tables(self, num_steps: int):
        t = self._tables.get(num_steps)
        if t is None:
            s = self.scheduler.set_timesteps(num_steps)
            t = (torch.from_numpy(s.timesteps).to(self.device), torch.from_numpy(s.sigmas).to(self.device),
                 s.init_noise_sigma)
            self._tables[num_steps] = t
        return t

    def set_scale(self, scale: float) -> None:
        """IPAdapter.set_scale / pipeline.set_scale (ip_adapter.py:179-182, custom_pipelines.py:17-20)."""
        for p in self.unet.attn_processors.values():
            if hasattr(p, "to_k_ip"):
                p.scale = scale

    def _buffers(self, n: int, h: int, w: int, L: int):
        key = (n, h, w, L)
        st = self._static.get(key)
        if st is None:
            cfg = self.unet.config
            dev = self.device
            st = {
                "latents": torch.empty((n, cfg.in_channels, h, w), dtype=torch.float16, device=dev),
                "model_in": torch.empty((2 * n, cfg.in_channels, h, w), dtype=torch.float16, device=dev),
                "ehs": torch.empty((2 * n, L, cfg.cross_attention_dim), dtype=torch.float16, device=dev),
                "text_embeds": torch.empty((2 * n, cfg.pooled_embed_dim), dtype=torch.float16, device=dev),
                "time_ids": torch.empty((2 * n, 6), dtype=torch.float32, device=dev),
                "step": torch.zeros((1,), dtype=torch.int32, device=dev),
            }
            self._static[key] = st
        return st

    def _step(self, st, timesteps, sigmas, guidance):
        noise = self.unet(st["model_in"], timesteps, st["ehs"], st["text_embeds"], st["time_ids"], step=st["step"])
        ops.euler_cfg_step(noise, st["latents"], st["model_in"], sigmas, st["step"], guidance)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, latents: torch.Tensor, prompt_embeds: torch.Tensor, negative_prompt_embeds: torch.Tensor,
            pooled: torch.Tensor, negative_pooled: torch.Tensor, time_ids: torch.Tensor, num_inference_steps: int,
            guidance_scale: float = 5.0, ip_scale: float = 1.0, control_guidance_start: float = 0.0,
            control_guidance_end: float = 1.0, stop_after: Optional[int] = None, start_step: int = 0) -> torch.Tensor:
        """latents [n,4,h,w] (already scaled by init_noise_sigma; host or device); embeds host or device tensors.
        Returns the latents after the last executed step as a new device tensor.  `stop_after` stops after step index
        k - 1 (PNS preview); `start_step` = k resumes a trajectory whose `latents` are the output of a `stop_after=k`
        call with the same schedule (two-phase PNS: preview all candidates, continue only the winner)."""
        if guidance_scale <= 1.0:
            raise IHError("the native loop implements the classifier-free-guidance path (guidance_scale > 1)")
        n, _, h, w = latents.shape
        L = prompt_embeds.shape[1]
        T = num_inference_steps
        timesteps, sigmas, _ = self.tables(T)
        st = self._buffers(n, h, w, L)
        # inputs -> static device buffers (H2D copies when the caller hands over pinned host tensors)
        st["latents"].copy_(latents, non_blocking=True)
        st["ehs"][:n].copy_(negative_prompt_embeds, non_blocking=True)        # CFG order [negative, positive], :296
        st["ehs"][n:].copy_(prompt_embeds, non_blocking=True)
        st["text_embeds"][:n].copy_(negative_pooled, non_blocking=True)
        st["text_embeds"][n:].copy_(pooled, non_blocking=True)
        st["time_ids"][:n].copy_(time_ids, non_blocking=True)
        st["time_ids"][n:].copy_(time_ids, non_blocking=True)
        if not 0 <= start_step <= T:
            raise IHError(f"start_step {start_step} outside the {T}-step schedule")
        st["step"].fill_(start_step)
        self.unet.prepare_conditioning(st["ehs"], st["text_embeds"], st["time_ids"])
        ops.scale_model_input(st["latents"], st["model_in"], sigmas, st["step"])   # :332-334 for the first step run

        steps = T if stop_after is None else min(stop_after, T)

        def scale_at(i: int) -> float:
            off = (i / T < control_guidance_start) or ((i + 1) / T > control_guidance_end)   # :326-329
            return 0.0 if off else float(ip_scale)

        def gkey(scale: float):
            return (n, h, w, L, T, float(guidance_scale), scale)

        if self.use_cuda_graph:
            missing = sorted({scale_at(i) for i in range(start_step, steps)} -
                             {k[-1] for k in self._graphs if k[:-1] == gkey(0.0)[:-1]})
            if missing:
                for sc in missing:
                    self.set_scale(sc)
                    self._graphs[gkey(sc)] = self._capture(st, timesteps, sigmas, guidance_scale)
                # the warm-up pass of a capture advances the state once: restore the initial state
                st["latents"].copy_(latents, non_blocking=True)
                st["step"].fill_(start_step)
                ops.scale_model_input(st["latents"], st["model_in"], sigmas, st["step"])
        for i in range(start_step, steps):
            sc = scale_at(i)
            if self.use_cuda_graph:
                self._graphs[gkey(sc)].replay()
            else:
                self.set_scale(sc)
                self._step(st, timesteps, sigmas, guidance_scale)
        return st["latents"].clone()

    def _capture(self, st, timesteps, sigmas, guidance):
        # warm-up on a side stream (sets kernel attributes, fills the TMA descriptor cache, sizes the allocator)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            st["step"].zero_()
            self._step(st, timesteps, sigmas, guidance)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        st["step"].zero_()
        g = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        with torch.cuda.graph(g):
            self._step(st, timesteps, sigmas, guidance)
        self.last_launches_per_step = ops.launch_count() - before
        return g
