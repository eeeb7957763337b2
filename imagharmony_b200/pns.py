"""Preference-guided noise selection (PNS) across the GPUs of one box.

The reference has no PNS code (README.md:27 and the inference stage of assets/1.png only: N candidate noises ->
preview denoising -> a judge scores the candidates -> the best noise is denoised to the end), so this driver is
specified by BASELINE.json's north_star: candidates are an embarrassingly-parallel shard over ranks (one process per
GPU, weights replicated, no data-path collective), followed by ONE all_gather of the fp32 scores (N floats: 128 B for
N = 32) and a local argmax, so every rank agrees on the winner.  Candidate latents come from per-seed CPU generators
(ip_adapter/utils.py:86-87 list-of-seeds semantics), which makes a candidate bit-identical whatever rank / batch
slot it lands in.

Parity status: unpinned by the reference (no implementation there); the shard / gather logic is covered by
world_size-2 gloo tests on CPU and the NCCL path by the 8-GPU bench.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch


def shard_seeds(seeds: Sequence[int], rank: int, world: int) -> List[int]:
    """Contiguous block partition; ranks past the end get an empty shard (a rank with 0 candidates is legal)."""
    n = len(seeds)
    per = (n + world - 1) // world
    return list(seeds[rank * per: min(n, (rank + 1) * per)])


class LinearProbeScorer:
    """Synthetic judge for environments without CLIP / VLM weights: a fixed random linear probe of the candidate
    latent, score_i = <w, x_i> / ||w||.  Deterministic in `seed`, identical on every rank.  On the GPU the probe is
    one launch of the library's weight-streaming small-M kernel (`ih_linear_small_f16`); CPU tensors (gloo tests) use
    a plain dot product.  `imagharmony_b200.clip.ClipScorer` is the real judge (image-text cosine of the CLIP towers)."""

    def __init__(self, numel: int, seed: int = 1234, device="cpu"):
        g = torch.Generator("cpu").manual_seed(seed)
        w = torch.randn(numel, generator=g)
        self.w = (w / w.norm()).to(device)
        self._w16 = None

    def describe(self) -> str:
        return "fixed random linear probe of the final latent (synthetic judge, no CLIP weights offline)"

    def __call__(self, latents: torch.Tensor) -> torch.Tensor:
        x = latents.reshape(latents.shape[0], -1)
        if x.is_cuda and x.dtype == torch.float16 and x.shape[0] <= 64:
            from . import ops
            if self._w16 is None or self._w16.device != x.device:
                # 8 output rows (16-byte rows for the kernel), row 0 carries the probe scaled into fp16's normal range
                self._w16 = torch.zeros((8, x.shape[1]), dtype=torch.float16, device=x.device)
                self._w16[0] = (self.w.to(x.device) * 64.0).half()
            return ops.linear_small(x.contiguous(), self._w16)[:, 0].float() / 64.0
        return x.float() @ self.w.to(latents.device)


@dataclass
class PNSResult:
    scores: torch.Tensor          # [N] fp32, candidate order = order of `seeds`
    best_index: int
    best_seed: int
    best_latents: Optional[torch.Tensor]   # final latents of the winner (on every rank when broadcast=True)


def gather_scores(local: torch.Tensor, counts: List[int], dist=None) -> torch.Tensor:
    """all_gather of per-rank score vectors of (possibly) unequal length -> [sum(counts)] in rank order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local.float().cpu()
    world = dist.get_world_size()
    width = max(max(counts), 1)
    buf = torch.full((width,), float("-inf"), dtype=torch.float32, device=local.device)
    buf[: local.numel()] = local.float()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)                      # the only collective of the path: world * width fp32 values
    return torch.cat([o[:c].cpu() for o, c in zip(out, counts)])


def pns_two_phase(run_preview: Callable[[List[int]], torch.Tensor], run_rest: Callable[[torch.Tensor], torch.Tensor],
                  seeds: Sequence[int], scorer: Callable, dist=None, max_batch: int = 4) -> "PNSResult":
    """The inference stage of assets/1.png: every candidate is denoised for a few PREVIEW steps, the judge scores the
    previews, and only the winner is denoised to the end.

    run_preview(list_of_seeds) -> preview latents [len, C, h, w] (e.g. DenoiseEngine.run(..., stop_after=k));
    run_rest(preview_latent [1, C, h, w]) -> final latents (DenoiseEngine.run(..., start_step=k)).
    Collectives: the all_gather of N fp32 scores and one broadcast of the winning preview latent (128 KiB at 1024^2);
    every rank then finishes the winner redundantly, so all ranks return the same `best_latents` without a third
    collective (the remaining steps of ONE trajectory do not shard)."""
    res = pns_select(run_preview, seeds, scorer, dist=dist, max_batch=max_batch, broadcast_winner=True)
    final = run_rest(res.best_latents.unsqueeze(0))
    return PNSResult(scores=res.scores, best_index=res.best_index, best_seed=res.best_seed, best_latents=final[0])


def pns_select(run_candidates: Callable[[List[int]], torch.Tensor], seeds: Sequence[int], scorer: Callable,
               dist=None, max_batch: int = 4, broadcast_winner: bool = True) -> PNSResult:
    """run_candidates(list_of_seeds) -> final (or preview) latents [len, C, h, w] for those seeds on this rank.

    Every rank: shard -> run its candidates in batches of `max_batch` -> score -> all_gather -> argmax."""
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_seeds(seeds, rank, world)
    counts = [len(shard_seeds(seeds, r, world)) for r in range(world)]
    lat_parts, score_parts = [], []
    for i in range(0, len(mine), max_batch):
        lat = run_candidates(mine[i:i + max_batch])
        lat_parts.append(lat)
        score_parts.append(scorer(lat))
    dev = lat_parts[0].device if lat_parts else (torch.device("cuda", torch.cuda.current_device())
                                                 if torch.cuda.is_available() and world > 1 and
                                                 dist.get_backend() == "nccl" else torch.device("cpu"))
    local = torch.cat(score_parts) if score_parts else torch.empty(0, device=dev)
    scores = gather_scores(local.to(dev), counts, dist)
    best = int(torch.argmax(scores).item())
    owner, off = 0, best
    for r, c in enumerate(counts):
        if off < c:
            owner = r
            break
        off -= c
    best_lat = None
    if broadcast_winner:
        if world == 1:
            best_lat = torch.cat(lat_parts)[off].clone()
        else:
            if rank == owner:
                best_lat = torch.cat(lat_parts)[off].contiguous().clone().to(torch.float16)
            if all(c > 0 for c in counts):
                shape = tuple(lat_parts[0].shape[1:])      # every rank holds candidates of the same shape: no exchange
            else:                                          # a rank without candidates must learn the latent shape
                st = (torch.tensor(list(best_lat.shape), device=dev, dtype=torch.int64) if rank == owner
                      else torch.zeros(3, device=dev, dtype=torch.int64))
                dist.broadcast(st, src=owner)
                shape = tuple(int(v) for v in st.tolist())
            if rank != owner:
                best_lat = torch.empty(shape, dtype=torch.float16, device=dev)
            dist.broadcast(best_lat, src=owner)       # 128 KiB at 1024^2: latency-bound, NVLink irrelevant
    return PNSResult(scores=scores, best_index=best, best_seed=int(seeds[best]), best_latents=best_lat)
