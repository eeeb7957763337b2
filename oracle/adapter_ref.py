"""ORACLE (test infrastructure only -- never imported by the product path).

Plain-PyTorch restatement of the reference-owned pieces of the hot path, SDXL case only (3-D hidden states, no
mask / spatial_norm / group_norm / norm_cross / residual, rescale 1.0).  Each function cites the reference lines it
follows.  Pinned against the real reference modules (loaded by file path from /root/reference in the build
container) by oracle/check_against_reference.py; the outputs of that run are committed under tests/golden/.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _split_heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, c = x.shape
    return x.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3)


def _merge_heads(x: torch.Tensor) -> torch.Tensor:
    b, h, n, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, n, h * d)


def _sdpa(q, k, v):
    """softmax(q k^T / sqrt(d)) v -- what F.scaled_dot_product_attention computes with no mask, p=0
    (attention_processor.py:312-314, 423-425, 440-442). Written out so the oracle does not depend on SDPA backends."""
    if q.is_cuda:
        # on a GPU the oracle doubles as the "reference GPU path" stand-in: call the very function the reference calls
        return F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    return torch.matmul(torch.softmax(s, dim=-1), v)


class SelfAttnProcessorRef(nn.Module):
    """AttnProcessor2_0.__call__ (attention_processor.py:258-332): q,k,v projections of the same tokens (or of the
    encoder tokens when given), SDPA, head merge, to_out[0] (+bias), dropout(0).  A parameter-less nn.Module like the
    reference's, so that ModuleList(unet.attn_processors.values()) keeps the checkpoint's index keys
    (ip_adapter.py:153-154)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = _split_heads(attn.to_q(hidden_states), attn.heads)        # :292, :305
        k = _split_heads(attn.to_k(ctx), attn.heads)                  # :299, :307
        v = _split_heads(attn.to_v(ctx), attn.heads)                  # :300, :308
        o = _merge_heads(_sdpa(q, k, v)).to(q.dtype)                  # :312-317
        return attn.to_out[1](attn.to_out[0](o))                      # :320-322


class IPAttnProcessorRef(nn.Module):
    """IPAttnProcessor2_0 (attention_processor.py:335-465).

    The last `num_tokens` encoder tokens are the image-prompt tokens (:402-406) and are *always* cut off the text
    context, even when skip=True (:430).  Active layers add  scale * SDPA(q, to_k_ip(ip), to_v_ip(ip))  (:432-450);
    the two softmaxes are independent ("decoupled").  `attn_map` (:443-444) is reproduced only on request because it
    has no consumer in the reference.
    """

    def __init__(self, hidden_size: int, cross_attention_dim: int, scale: float = 1.0, num_tokens: int = 4,
                 skip: bool = False, keep_attn_map: bool = False):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        self.num_tokens = num_tokens
        self.skip = skip
        self.keep_attn_map = keep_attn_map
        self.to_k_ip = nn.Linear(cross_attention_dim, hidden_size, bias=False)   # :361
        self.to_v_ip = nn.Linear(cross_attention_dim, hidden_size, bias=False)   # :362

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        q = _split_heads(attn.to_q(hidden_states), attn.heads)                    # :396, :416
        n_text = encoder_hidden_states.shape[1] - self.num_tokens                 # :402
        text, ip = encoder_hidden_states[:, :n_text], encoder_hidden_states[:, n_text:]   # :403-406
        k = _split_heads(attn.to_k(text), attn.heads)                             # :410, :418
        v = _split_heads(attn.to_v(text), attn.heads)                             # :411, :419
        o = _merge_heads(_sdpa(q, k, v)).to(q.dtype)                              # :423-428
        if not self.skip:                                                         # :430
            k_ip = _split_heads(self.to_k_ip(ip), attn.heads)                     # :432, :435
            v_ip = _split_heads(self.to_v_ip(ip), attn.heads)                     # :433, :436
            o_ip = _merge_heads(_sdpa(q, k_ip, v_ip)).to(q.dtype)                 # :440-448
            if self.keep_attn_map:
                # note the reference's operator precedence: softmax is applied to k_ip^T (over the token axis), unscaled
                self.attn_map = torch.matmul(q, k_ip.transpose(-2, -1).softmax(dim=-1))   # :443-444
            o = o + self.scale * o_ip                                             # :450
        return attn.to_out[1](attn.to_out[0](o))                                  # :453-455


class CrossAttentionHARef(nn.Module):
    """Cross_Attention (attention_processor.py:12-56): biased q/k/v projections, head_dim = query_dim // heads,
    scores divided by sqrt(head_dim), value width `value_dim` per head, biased out_proj."""

    def __init__(self, query_dim: int, context_dim: int, heads: int, value_dim: int):
        super().__init__()
        self.heads = heads
        self.head_dim = query_dim // heads                                        # :22
        self.value_dim = value_dim
        self.to_q = nn.Linear(query_dim, heads * self.head_dim)                   # :28
        self.to_k = nn.Linear(context_dim, heads * self.head_dim)                 # :29
        self.to_v = nn.Linear(context_dim, heads * value_dim)                     # :30
        self.out_proj = nn.Linear(heads * value_dim, heads * value_dim)           # :33

    def forward(self, query_input, context_input):
        b = query_input.shape[0]                                                  # :37 (batch taken from the query!)
        q = self.to_q(query_input).reshape(b, -1, self.heads, self.head_dim).transpose(1, 2)
        k = self.to_k(context_input).reshape(b, -1, self.heads, self.head_dim).transpose(1, 2)
        v = self.to_v(context_input).reshape(b, -1, self.heads, self.value_dim).transpose(1, 2)
        p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.head_dim), dim=-1)   # :45-46
        o = torch.matmul(p, v).transpose(1, 2).reshape(b, -1, self.heads * self.value_dim)           # :49-52
        return self.out_proj(o)                                                   # :55


class HarmonyAttentionRef(nn.Module):
    """HarmonyAttention with fusion_method="cross_attention" (train.py:188-266): fc1 -> split into reshape_blocks
    query tokens -> cross attention over the auxiliary text tokens -> flatten -> LayerNorm -> fc2 -> * scale.
    The caller adds the result to the CLIP image embedding (ip_adapter.py:172-173)."""

    def __init__(self, image_hidden_size=1280, text_context_dim=2048, inter_dim=2560, cross_heads=8,
                 reshape_blocks=8, cross_value_dim=64, scale=1.0):
        super().__init__()
        self.scale = scale
        self.reshape_blocks = reshape_blocks
        self.cross_query_dim = inter_dim // reshape_blocks                        # :202
        self.fc1 = nn.Linear(image_hidden_size, inter_dim)                        # :208
        self.fusion_text_image = CrossAttentionHARef(self.cross_query_dim, text_context_dim, cross_heads,
                                                     cross_value_dim)             # :212-217
        flat = cross_value_dim * cross_heads * reshape_blocks                     # :237
        self.ln = nn.LayerNorm(flat)                                              # :238
        self.fc2 = nn.Linear(flat, image_hidden_size)                             # :239

    def forward(self, text_embeds, image_embeds):
        b = image_embeds.shape[0]
        x = self.fc1(image_embeds).reshape(b, self.reshape_blocks, self.cross_query_dim)   # :254-255
        a = self.fusion_text_image(x, text_embeds).reshape(b, -1)                 # :259-262
        return self.fc2(self.ln(a)) * self.scale                                  # :263-264


class ImageProjRef(nn.Module):
    """ImageProjModel (ip_adapter.py:28-48): Linear -> [.., tokens, cross_dim] -> LayerNorm."""

    def __init__(self, cross_attention_dim=2048, clip_embeddings_dim=1280, clip_extra_context_tokens=4):
        super().__init__()
        self.cross_attention_dim = cross_attention_dim
        self.clip_extra_context_tokens = clip_extra_context_tokens
        self.proj = nn.Linear(clip_embeddings_dim, clip_extra_context_tokens * cross_attention_dim)
        self.norm = nn.LayerNorm(cross_attention_dim)

    def forward(self, image_embeds):
        t = self.proj(image_embeds).reshape(-1, self.clip_extra_context_tokens, self.cross_attention_dim)
        return self.norm(t)


class PerceiverAttentionRef(nn.Module):
    """PerceiverAttention (resampler.py:34-78): q from the latents, k/v from cat(x, latents); q and k are each
    pre-scaled by dim_head**-0.25; softmax in fp32."""

    def __init__(self, dim: int, dim_head: int, heads: int):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, latents):
        x = self.norm1(x)
        latents = self.norm2(latents)
        q = _split_heads(self.to_q(latents), self.heads)
        k, v = self.to_kv(torch.cat([x, latents], dim=1)).chunk(2, dim=-1)
        k, v = _split_heads(k, self.heads), _split_heads(v, self.heads)
        s = self.dim_head ** -0.25
        w = torch.matmul(q * s, (k * s).transpose(-1, -2))                        # :71-72
        w = torch.softmax(w.float(), dim=-1).to(w.dtype)                          # :73
        return self.to_out(_merge_heads(torch.matmul(w, v)))                      # :74-78


class ResamplerRef(nn.Module):
    """Resampler (resampler.py:81-147), parameter names as in the reference state dict."""

    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024,
                 ff_mult=4, max_seq_len=257, apply_pos_emb=False, num_latents_mean_pooled=0):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, embedding_dim) if apply_pos_emb else None
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.num_latents_mean_pooled = num_latents_mean_pooled
        if num_latents_mean_pooled > 0:
            # index 2 of the reference Sequential is a parameter-free Rearrange
            self.to_latents_from_mean_pooled_seq = nn.Sequential(nn.LayerNorm(dim),
                                                                 nn.Linear(dim, dim * num_latents_mean_pooled))
        else:
            self.to_latents_from_mean_pooled_seq = None
        inner = int(dim * ff_mult)
        self.layers = nn.ModuleList([
            nn.ModuleList([PerceiverAttentionRef(dim, dim_head, heads),
                           nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), nn.GELU(),
                                         nn.Linear(inner, dim, bias=False))])
            for _ in range(depth)])

    def forward(self, x):
        if self.pos_emb is not None:
            x = x + self.pos_emb(torch.arange(x.shape[1], device=x.device))      # :128-131
        latents = self.latents.repeat(x.shape[0], 1, 1)                           # :133
        x = self.proj_in(x)                                                       # :135
        if self.to_latents_from_mean_pooled_seq is not None:
            pooled = self.to_latents_from_mean_pooled_seq(x.mean(dim=1))          # :137-139 (all-ones mask)
            pooled = pooled.reshape(x.shape[0], self.num_latents_mean_pooled, -1)
            latents = torch.cat([pooled, latents], dim=1)                         # :140
        for attn, ff in self.layers:
            latents = attn(x, latents) + latents                                  # :143
            latents = ff(latents) + latents                                       # :144
        return self.norm_out(self.proj_out(latents))                              # :146-147


def install_processors(unet, cfg, scale: float = 1.0, dtype=torch.float32):
    """IPAdapter.set_ip_adapter (ip_adapter.py:99-125): attn1 -> self-attention processor; attn2 -> IP processor,
    active only where the processor name contains `cfg.ip_target_substring` (:117), skip=True elsewhere (:121-123)."""
    procs = {}
    boc = cfg.block_out_channels
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            hidden = boc[-1]
        elif name.startswith("up_blocks"):
            hidden = list(reversed(boc))[int(name[len("up_blocks.")])]
        else:
            hidden = boc[int(name[len("down_blocks.")])]
        if name.endswith("attn1.processor"):
            procs[name] = SelfAttnProcessorRef()
        else:
            procs[name] = IPAttnProcessorRef(hidden, cfg.cross_attention_dim, scale=scale,
                                             num_tokens=cfg.num_ip_tokens,
                                             skip=cfg.ip_target_substring not in name).to(dtype)
    unet.set_attn_processor(procs)
    return procs
