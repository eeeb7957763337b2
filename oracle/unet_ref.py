"""ORACLE (test infrastructure only -- never imported by the product path).

CPU / plain-PyTorch restatement of the SDXL `UNet2DConditionModel` forward that the reference calls at
ip_adapter/custom_pipelines.py:338-345 (and train.py:310).  The arithmetic lives in the third-party dependency
`diffusers==0.30.0` (/root/reference/requirements.txt:25), which is NOT vendored under /root/reference and is not
installable here; this file restates its published algorithm for the SDXL-base configuration (SURVEY.md appendix
A.1/A.2): ResnetBlock2D, Transformer2DModel (use_linear_projection), BasicTransformerBlock, GEGLU feed-forward,
Downsample2D / Upsample2D, text_time additional embedding.  Module and parameter names mirror diffusers' state-dict
keys exactly so real SDXL weights / the reference's ip_adapter.bin would load without a key map.

Parity status: **unpinned** for the [3P] UNet arithmetic (the reference holds no golden vectors for it,
SURVEY.md section 8c); the attention-processor boundary *is* pinned against the reference's own
ip_adapter/attention_processor.py (see oracle/check_against_reference.py and tests/golden/).

Attention goes through the same processor protocol the reference plugs into:
    attn.processor(attn, hidden_states, encoder_hidden_states=..., attention_mask=None)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from imagharmony_b200.config import UNetConfig


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0) -> [cos | sin], fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.reshape(-1, 1).float() * freqs.reshape(1, -1)
    return torch.cat([ang.cos(), ang.sin()], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        skip = x if self.conv_shortcut is None else self.conv_shortcut(x)
        return skip + h


class Attention(nn.Module):
    """The slice of diffusers' Attention the reference processors touch (attention_processor.py:374-463)."""

    def __init__(self, query_dim: int, heads: int, cross_attention_dim: Optional[int] = None):
        super().__init__()
        kv_dim = cross_attention_dim or query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(kv_dim, query_dim, bias=False)
        self.to_v = nn.Linear(kv_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim), nn.Dropout(0.0)])
        # attributes the processors read; all inert in SDXL blocks
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = False
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = None

    def forward(self, hidden_states, encoder_hidden_states=None):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=None)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        value, gate = self.proj(x).chunk(2, dim=-1)
        return value * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, cross_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h, ehs):
        h = h + self.attn1(self.norm1(h))
        h = h + self.attn2(self.norm2(h), ehs)
        h = h + self.ff(self.norm3(h))
        return h


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, depth: int, cross_dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ehs):
        b, c, hh, ww = x.shape
        h = self.norm(x).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ehs)
        h = self.proj_out(h).reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        return h + x


class Resample(nn.Module):
    def __init__(self, ch: int, up: bool):
        super().__init__()
        self.up = up
        self.conv = nn.Conv2d(ch, ch, 3, stride=1 if up else 2, padding=1)

    def forward(self, x):
        if self.up:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin: int, cout: int, depth: int, add_down: bool):
        super().__init__()
        temb = cfg.time_embed_dim
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, temb, cfg.norm_num_groups)
                                      for j in range(cfg.layers_per_block)])
        if depth > 0:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, cfg.heads(cout), depth, cfg.cross_attention_dim,
                                                                cfg.norm_num_groups)
                                             for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        if add_down:
            self.downsamplers = nn.ModuleList([Resample(cout, up=False)])
        else:
            self.downsamplers = None

    def forward(self, x, emb, ehs, skips: List[torch.Tensor]):
        for j, res in enumerate(self.resnets):
            x = res(x, emb)
            if self.attentions is not None:
                x = self.attentions[j](x, ehs)
            skips.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            skips.append(x)
        return x


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, ch: int, depth: int):
        super().__init__()
        temb = cfg.time_embed_dim
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, cfg.norm_num_groups) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, cfg.heads(ch), depth, cfg.cross_attention_dim,
                                                            cfg.norm_num_groups)])

    def forward(self, x, emb, ehs):
        x = self.resnets[0](x, emb)
        x = self.attentions[0](x, ehs)
        return self.resnets[1](x, emb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, prev_out: int, skip_ch: int, cout: int, depth: int, add_up: bool):
        super().__init__()
        temb = cfg.time_embed_dim
        n = cfg.layers_per_block + 1
        res = []
        for j in range(n):
            res_skip = skip_ch if j == n - 1 else cout
            res_in = prev_out if j == 0 else cout
            res.append(ResnetBlock2D(res_in + res_skip, cout, temb, cfg.norm_num_groups))
        self.resnets = nn.ModuleList(res)
        if depth > 0:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, cfg.heads(cout), depth, cfg.cross_attention_dim,
                                                                cfg.norm_num_groups) for _ in range(n)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Resample(cout, up=True)]) if add_up else None

    def forward(self, x, emb, ehs, skips: List[torch.Tensor]):
        for j, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, emb)
            if self.attentions is not None:
                x = self.attentions[j](x, ehs)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNetRef(nn.Module):
    """SDXL-shaped UNet2DConditionModel restatement (fp32 on CPU by default)."""

    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.config = cfg
        boc = cfg.block_out_channels
        tl = cfg.transformer_layers_per_block
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        self.add_embedding = TimestepEmbedding(cfg.add_embed_in, cfg.time_embed_dim)
        downs = []
        ch = boc[0]
        for i, co in enumerate(boc):
            downs.append(DownBlock(cfg, ch, co, tl[i], add_down=i < len(boc) - 1))
            ch = co
        self.down_blocks = nn.ModuleList(downs)
        rev = list(reversed(boc))
        rtl = list(reversed(tl))
        ups = []
        prev = rev[0]
        for i, co in enumerate(rev):
            skip_ch = rev[min(i + 1, len(rev) - 1)]
            ups.append(UpBlock(cfg, prev, skip_ch, co, rtl[i], add_up=i < len(rev) - 1))
            prev = co
        self.up_blocks = nn.ModuleList(ups)
        self.mid_block = MidBlock(cfg, boc[-1], tl[-1])
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    # --- the slice of the diffusers API the reference adapter uses (ip_adapter.py:102,125,153,180) ---
    @property
    def attn_processors(self) -> Dict[str, object]:
        out = {}
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                out[f"{name}.processor"] = m.processor
        return out

    def set_attn_processor(self, procs: Dict[str, object]) -> None:
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                m.processor = procs[f"{name}.processor"]

    def forward(self, sample, timestep, encoder_hidden_states, text_embeds, time_ids):
        cfg = self.config
        dt = sample.dtype
        b = sample.shape[0]
        if torch.is_tensor(timestep):
            t = timestep.to(device=sample.device, dtype=torch.float32).reshape(-1).expand(b)
        else:
            t = torch.full((b,), float(timestep), dtype=torch.float32, device=sample.device)
        emb = self.time_embedding(sinusoidal_embedding(t, cfg.block_out_channels[0]).to(dt))
        tid = sinusoidal_embedding(time_ids.reshape(-1), cfg.addition_time_embed_dim).reshape(b, -1)
        aug = self.add_embedding(torch.cat([text_embeds, tid.to(text_embeds.dtype)], dim=-1).to(dt))
        emb = emb + aug
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x = blk(x, emb, encoder_hidden_states, skips)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, emb, encoder_hidden_states, skips)
        return self.conv_out(F.silu(self.conv_norm_out(x)))
