"""Once-per-generate() adapter modules on the sm_100a kernels, with the reference's class names, constructor
arguments and state-dict keys:

  ImageProjModel      ip_adapter/ip_adapter.py:28-48      Linear 1280 -> 4*2048, reshape, LayerNorm(2048)
  Cross_Attention     ip_adapter/attention_processor.py:12-56   (biased q/k/v, head_dim = query_dim // heads, v_dim)
  HarmonyAttention    train.py:188-266  (fusion_method="cross_attention" only -- the other variants are marked TODO
                                         in the reference and crash at the shipped hyper-parameters)
  Resampler           ip_adapter/resampler.py:81-147  (+ PerceiverAttention :34-78, FeedForward :13-20)

nn.Linear / nn.LayerNorm / nn.Embedding are parameter containers; the arithmetic is ops.* (C-ABI kernels).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from ._lib import IHError


class ImageProjModel(nn.Module):
    def __init__(self, cross_attention_dim=1024, clip_embeddings_dim=1024, clip_extra_context_tokens=4):
        super().__init__()
        self.generator = None
        self.cross_attention_dim = cross_attention_dim
        self.clip_extra_context_tokens = clip_extra_context_tokens
        self.proj = nn.Linear(clip_embeddings_dim, clip_extra_context_tokens * cross_attention_dim)
        self.norm = nn.LayerNorm(cross_attention_dim)

    def forward(self, image_embeds: torch.Tensor) -> torch.Tensor:
        x = image_embeds.reshape(-1, image_embeds.shape[-1]).contiguous()
        t = ops.linear_small(x, self.proj.weight, self.proj.bias)                       # ip_adapter.py:44
        t = t.reshape(-1, self.clip_extra_context_tokens, self.cross_attention_dim)     # :45-46
        return ops.layernorm(t, self.norm.weight, self.norm.bias, self.norm.eps)        # :47


class Cross_Attention(nn.Module):
    def __init__(self, query_dim, context_dim, heads=8, value_dim=None, out_dim=None):
        super().__init__()
        self.query_dim = query_dim
        self.heads = heads
        self.head_dim = query_dim // heads                                              # attention_processor.py:22
        self.scale = math.sqrt(self.head_dim)                                           # :23 (a divisor)
        self.value_dim = value_dim if value_dim is not None else self.head_dim
        self.out_dim = out_dim if out_dim is not None else heads * self.value_dim
        self.to_q = nn.Linear(query_dim, heads * self.head_dim)
        self.to_k = nn.Linear(context_dim, heads * self.head_dim)
        self.to_v = nn.Linear(context_dim, heads * self.value_dim)
        self.out_proj = nn.Linear(heads * self.value_dim, self.out_dim)

    def forward(self, query_input: torch.Tensor, context_input: torch.Tensor) -> torch.Tensor:
        B, Nq, _ = query_input.shape                                                    # :37 batch from the query
        ctx = context_input.reshape(B, -1, context_input.shape[-1])                     # text batch folds into keys
        Nk = ctx.shape[1]
        q = ops.linear(query_input.reshape(B * Nq, -1).contiguous(), self.to_q.weight, self.to_q.bias)
        k = ops.linear(ctx.reshape(B * Nk, -1).contiguous(), self.to_k.weight, self.to_k.bias)
        v = ops.linear(ctx.reshape(B * Nk, -1).contiguous(), self.to_v.weight, self.to_v.bias)
        o = ops.attention_small(q, k, v, B, self.heads, Nq, Nk, self.head_dim, self.value_dim, self.scale)   # :45-52
        out = ops.linear(o, self.out_proj.weight, self.out_proj.bias)                   # :55
        return out.reshape(B, Nq, self.out_dim)


class HarmonyAttention(nn.Module):
    def __init__(self, image_hidden_size=1280, text_context_dim=2048, inter_dim=2560, cross_heads=10,
                 reshape_blocks=8, cross_value_dim=64, scale=1.0, fusion_method="cross_attention"):
        super().__init__()
        if fusion_method != "cross_attention":
            raise IHError(f"fusion_method={fusion_method!r}: only 'cross_attention' is supported (the reference's "
                          "qformer / mlp / gated-attention variants are TODO stubs that fail at the shipped sizes)")
        self.scale = scale
        self.reshape_blocks = reshape_blocks
        self.cross_query_dim = inter_dim // reshape_blocks
        self.fusion_method = fusion_method
        self.image_hidden_size = image_hidden_size
        self.text_context_dim = text_context_dim
        self.fc1 = nn.Linear(image_hidden_size, inter_dim)
        self.fusion_text_image = Cross_Attention(query_dim=self.cross_query_dim, context_dim=text_context_dim,
                                                 heads=cross_heads, value_dim=cross_value_dim)
        flattened_dim = cross_value_dim * cross_heads * reshape_blocks
        self.ln = nn.LayerNorm(flattened_dim)
        self.fc2 = nn.Linear(flattened_dim, image_hidden_size)

    def forward(self, text_embeds: torch.Tensor, image_embeds: torch.Tensor, add_to: torch.Tensor = None):
        """-> fc2(LN(attn)) * scale  (train.py:243-266; no prints).  With `add_to` the caller's
        `clip_image_embeds + output` (ip_adapter.py:173) is fused into the last kernel."""
        B = image_embeds.shape[0]
        x = ops.linear_small(image_embeds.contiguous(), self.fc1.weight, self.fc1.bias)             # :254
        x = x.reshape(B, self.reshape_blocks, self.cross_query_dim)                                 # :255
        # encode_prompt(extra_text, num_images_per_prompt=n) lays the rows out [a,a,..,b,b,..]; the reference's
        # Cross_Attention folds each image's n copies into its key axis (view(B, -1, D), attention_processor.py:37-41).
        # Repeated keys do not change a softmax-weighted mean (C.11), so ONE copy per image group is taken instead.
        rows = text_embeds.shape[0]
        if rows % B != 0:
            raise IHError(f"HarmonyAttention: {rows} auxiliary-text rows cannot be grouped over {B} image embeddings")
        attended = self.fusion_text_image(x, text_embeds[:: rows // B].contiguous())
        a = ops.layernorm(attended.reshape(B, -1), self.ln.weight, self.ln.bias, self.ln.eps)       # :262-263
        return ops.linear_small(a, self.fc2.weight, self.fc2.bias, out_scale=float(self.scale), addend=add_to)  # :264


ComposedAttention = HarmonyAttention   # demo.py:11,56-64 imports this name from tutorial_train_sdxl_ori


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        if dim_head != 64:
            raise IHError("PerceiverAttention: the tcgen05 attention kernel is specialised for dim_head = 64")
        self.dim_head, self.heads = dim_head, heads
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x: torch.Tensor, latents: torch.Tensor, residual: torch.Tensor = None) -> torch.Tensor:
        B, n1, D = x.shape
        n2 = latents.shape[1]
        inner = self.dim_head * self.heads
        xn = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        ln = ops.layernorm(latents, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        q = ops.linear(ln.reshape(B * n2, D), self.to_q.weight)
        kv = torch.empty((B, n1 + n2, 2 * inner), dtype=torch.float16, device=x.device)   # cat((x, latents)) -> to_kv
        for b in range(B):
            ops.linear(xn[b], self.to_kv.weight, out=kv[b, :n1])
            ops.linear(ln[b], self.to_kv.weight, out=kv[b, n1:])
        kv2 = kv.reshape(B * (n1 + n2), 2 * inner)
        # (q * s)(k * s)^T with s = 64^-0.25 is q k^T / 8: exactly the kernel's softmax scale (resampler.py:71-73)
        o = ops.attention(q, kv2[:, :inner], kv2[:, inner:], B, self.heads, n2, n1 + n2)
        res = None if residual is None else residual.reshape(B * n2, D)
        return ops.linear(o, self.to_out.weight, residual=res).reshape(B, n2, D)


class _FeedForward(nn.Sequential):
    """LayerNorm, Linear(no bias), GELU, Linear(no bias) -- same Sequential indices as resampler.py:13-20."""

    def __init__(self, dim, mult=4):
        inner = int(dim * mult)
        super().__init__(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), nn.GELU(),
                         nn.Linear(inner, dim, bias=False))

    def forward(self, x: torch.Tensor, residual: torch.Tensor = None) -> torch.Tensor:
        B, n, D = x.shape
        h = ops.layernorm(x, self[0].weight, self[0].bias, self[0].eps)
        h = ops.linear(h.reshape(B * n, D), self[1].weight, gelu=True)
        res = None if residual is None else residual.reshape(B * n, D)
        return ops.linear(h, self[3].weight, residual=res).reshape(B, n, D)


class Resampler(nn.Module):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024,
                 ff_mult=4, max_seq_len: int = 257, apply_pos_emb: bool = False, num_latents_mean_pooled: int = 0):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, embedding_dim) if apply_pos_emb else None
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.num_latents_mean_pooled = num_latents_mean_pooled
        self.to_latents_from_mean_pooled_seq = (
            nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim * num_latents_mean_pooled))
            if num_latents_mean_pooled > 0 else None)
        self.layers = nn.ModuleList([nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads),
                                                    _FeedForward(dim=dim, mult=ff_mult)]) for _ in range(depth)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, n, E = x.shape
        x = x.contiguous()
        if self.pos_emb is not None:
            x = ops.add_bcast(x, self.pos_emb.weight[:n].contiguous())                            # :128-131
        dim = self.proj_in.weight.shape[0]
        x = ops.linear(x.reshape(B * n, E), self.proj_in.weight, self.proj_in.bias).reshape(B, n, dim)   # :135
        latents = self.latents.detach().to(torch.float16).repeat(B, 1, 1).contiguous()            # :133
        if self.to_latents_from_mean_pooled_seq is not None:
            pooled = ops.mean_tokens(x)                                                           # :137-138
            ln, lin = self.to_latents_from_mean_pooled_seq[0], self.to_latents_from_mean_pooled_seq[1]
            pooled = ops.layernorm(pooled, ln.weight, ln.bias, ln.eps)
            pooled = ops.linear_small(pooled, lin.weight, lin.bias).reshape(B, self.num_latents_mean_pooled, dim)
            latents = torch.cat((pooled, latents), dim=-2).contiguous()                           # :140 (host glue)
        for attn, ff in self.layers:
            latents = attn(x, latents, residual=latents)                                          # :143
            latents = ff(latents, residual=latents)                                               # :144
        Q = latents.shape[1]
        out = ops.linear(latents.reshape(B * Q, dim), self.proj_out.weight, self.proj_out.bias)   # :146
        return ops.layernorm(out.reshape(B, Q, -1), self.norm_out.weight, self.norm_out.bias, self.norm_out.eps)
