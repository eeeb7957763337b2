"""CLIP towers on the sm_100a kernel library (scope row f2: the conditioning encoders and the PNS judge).

The reference loads `CLIPVisionModelWithProjection` (ViT-bigG/14 for SDXL; ip_adapter.py:81-84) and calls it at
:163-164 / :404-412 (`image_embeds`, or `hidden_states[-2]` for the Plus variant), and reaches the two SDXL text
encoders (`CLIPTextModel` = CLIP ViT-L, `CLIPTextModelWithProjection` = OpenCLIP bigG) through
`pipe.encode_prompt` (:292-297, :314-319).  Those classes live in `transformers` ([3P], requirements.txt:145); here
their forward pass is restated on the C-ABI kernels:

    embeddings   text: ih_embed_tokens_f16 (token + position);  vision: patch conv as ONE tcgen05 GEMM per image
                 (patch rows @ flattened conv weight, position embedding fused as the residual operand)
    layer        LayerNorm -> q|k|v GEMM (one [3C, C] weight, bias) -> ih_attention_generic_f16 (causal for text, head_dim
                 64 / 80 / 104) -> out-proj GEMM (+bias +residual) -> LayerNorm -> fc1 GEMM (+bias, GELU or quick-GELU
                 epilogue) -> fc2 GEMM (+bias +residual)
    heads        text: final LayerNorm on the EOS rows -> text_projection; vision: post LayerNorm on the class token ->
                 visual_projection

State-dict keys are the `transformers` ones (`text_model.encoder.layers.N.self_attn.q_proj.weight`, ...), so a
checkpoint folder of the reference loads unchanged.  `ClipScorer` is the PNS judge BASELINE.json names ("allgather of
CLIP scores"): cosine of the bigG image embedding of a decoded candidate and the bigG text embedding of the prompt.

Parity: pinned against the `transformers` classes themselves (installed in this image) on random-init miniature and
full-size configurations -- tests/test_clip_cpu.py (wiring, stand-in ops) and tests/test_clip_gpu.py (kernels).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from ._lib import IHError

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # [3P] OPENAI_CLIP_MEAN / STD (CLIPImageProcessor defaults)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass
class ClipTowerConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    hidden_act: str = "gelu"                # "gelu" (erf) | "quick_gelu"
    layer_norm_eps: float = 1e-5
    projection_dim: Optional[int] = None
    # text
    vocab_size: int = 49408
    max_position_embeddings: int = 77
    eos_token_id: int = 2
    # vision
    image_size: int = 224
    patch_size: int = 14
    num_channels: int = 3

    @classmethod
    def from_hf(cls, c) -> "ClipTowerConfig":
        kw = dict(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                  num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                  hidden_act=c.hidden_act, layer_norm_eps=c.layer_norm_eps,
                  projection_dim=getattr(c, "projection_dim", None))
        for k in ("vocab_size", "max_position_embeddings", "eos_token_id", "image_size", "patch_size", "num_channels"):
            if getattr(c, k, None) is not None:
                kw[k] = getattr(c, k)
        return cls(**kw)


def tower_param_shapes(cfg: "ClipTowerConfig", kind: str) -> Dict[str, tuple]:
    """{transformers state-dict key: shape} of a CLIP text (`kind="text"`) or vision tower with projection head -- what
    `ClipTextTower` / `ClipVisionTower` consume; lets benchmarks build random-init towers straight on the GPU (no weights
    exist offline, and constructing a 1.8 B parameter `transformers` module on the host takes a minute)."""
    C, F, P = cfg.hidden_size, cfg.intermediate_size, cfg.projection_dim
    pre = "text_model." if kind == "text" else "vision_model."
    sh: Dict[str, tuple] = {}
    if kind == "text":
        sh[pre + "embeddings.token_embedding.weight"] = (cfg.vocab_size, C)
        sh[pre + "embeddings.position_embedding.weight"] = (cfg.max_position_embeddings, C)
    else:
        g = cfg.image_size // cfg.patch_size
        sh[pre + "embeddings.class_embedding"] = (C,)
        sh[pre + "embeddings.patch_embedding.weight"] = (C, cfg.num_channels, cfg.patch_size, cfg.patch_size)
        sh[pre + "embeddings.position_embedding.weight"] = (g * g + 1, C)
        sh[pre + "pre_layrnorm.weight"] = sh[pre + "pre_layrnorm.bias"] = (C,)
    for i in range(cfg.num_hidden_layers):
        p = f"{pre}encoder.layers.{i}."
        for n in "qkv":
            sh[p + f"self_attn.{n}_proj.weight"], sh[p + f"self_attn.{n}_proj.bias"] = (C, C), (C,)
        sh[p + "self_attn.out_proj.weight"], sh[p + "self_attn.out_proj.bias"] = (C, C), (C,)
        sh[p + "layer_norm1.weight"] = sh[p + "layer_norm1.bias"] = (C,)
        sh[p + "layer_norm2.weight"] = sh[p + "layer_norm2.bias"] = (C,)
        sh[p + "mlp.fc1.weight"], sh[p + "mlp.fc1.bias"] = (F, C), (F,)
        sh[p + "mlp.fc2.weight"], sh[p + "mlp.fc2.bias"] = (C, F), (C,)
    if kind == "text":
        sh[pre + "final_layer_norm.weight"] = sh[pre + "final_layer_norm.bias"] = (C,)
        if P:
            sh["text_projection.weight"] = (P, C)
    else:
        sh[pre + "post_layernorm.weight"] = sh[pre + "post_layernorm.bias"] = (C,)
        if P:
            sh["visual_projection.weight"] = (P, C)
    return sh


class _Layer:
    __slots__ = ("ln1", "ln2", "w_qkv", "b_qkv", "w_o", "b_o", "w_fc1", "b_fc1", "w_fc2", "b_fc2")


_DTYPE = [torch.float16]     # parameter / activation dtype; the CPU wiring tests switch it to fp32 (stand-in ops)


def _dev16(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=_DTYPE[0]).contiguous()


class _ClipEncoder:
    """The shared pre-LayerNorm transformer stack ([3P] CLIPEncoder / CLIPEncoderLayer)."""

    def __init__(self, cfg: ClipTowerConfig, sd: Dict[str, torch.Tensor], prefix: str, device):
        if cfg.hidden_act not in ("gelu", "quick_gelu"):
            raise IHError(f"CLIP activation {cfg.hidden_act!r} is not supported (gelu | quick_gelu)")
        C, H = cfg.hidden_size, cfg.num_attention_heads
        if C % H or (C // H) % 8:
            raise IHError(f"CLIP head_dim {C}/{H} must be a multiple of 8")
        self.cfg, self.device = cfg, device
        self.head_dim = C // H
        self.layers: List[_Layer] = []
        for i in range(cfg.num_hidden_layers):
            p = f"{prefix}encoder.layers.{i}."
            L = _Layer()
            L.ln1 = (_dev16(sd[p + "layer_norm1.weight"], device), _dev16(sd[p + "layer_norm1.bias"], device))
            L.ln2 = (_dev16(sd[p + "layer_norm2.weight"], device), _dev16(sd[p + "layer_norm2.bias"], device))
            L.w_qkv = _dev16(torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0), device)
            L.b_qkv = _dev16(torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0), device)
            L.w_o, L.b_o = _dev16(sd[p + "self_attn.out_proj.weight"], device), _dev16(sd[p + "self_attn.out_proj.bias"], device)
            L.w_fc1, L.b_fc1 = _dev16(sd[p + "mlp.fc1.weight"], device), _dev16(sd[p + "mlp.fc1.bias"], device)
            L.w_fc2, L.b_fc2 = _dev16(sd[p + "mlp.fc2.weight"], device), _dev16(sd[p + "mlp.fc2.bias"], device)
            self.layers.append(L)

    def layer(self, h: torch.Tensor, L: _Layer, B: int, N: int, causal: bool) -> torch.Tensor:
        cfg = self.cfg
        C, H, hd = cfg.hidden_size, cfg.num_attention_heads, self.head_dim
        n = ops.layernorm(h, L.ln1[0], L.ln1[1], cfg.layer_norm_eps)
        qkv = ops.linear(n, L.w_qkv, L.b_qkv)
        o = ops.attention_generic(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, hd, hd, hd ** -0.5, causal)
        h = ops.linear(o, L.w_o, L.b_o, residual=h)
        n = ops.layernorm(h, L.ln2[0], L.ln2[1], cfg.layer_norm_eps)
        f = ops.linear(n, L.w_fc1, L.b_fc1, gelu=cfg.hidden_act == "gelu", quick_gelu=cfg.hidden_act == "quick_gelu")
        return ops.linear(f, L.w_fc2, L.b_fc2, residual=h)

    def run(self, h: torch.Tensor, B: int, N: int, causal: bool, keep: Sequence[int] = ()) -> Dict[int, torch.Tensor]:
        """-> {index: hidden state} for the requested `hidden_states` indices (0 = embeddings, i = after layer i; negative
        indices count from the end like the HF tuple) plus the last one under key `len(layers)`."""
        nl = len(self.layers)
        want = {(i if i >= 0 else nl + 1 + i) for i in keep} | {nl}
        out = {}
        if 0 in want:
            out[0] = h
        for i, L in enumerate(self.layers):
            h = self.layer(h, L, B, N, causal)
            if i + 1 in want:
                out[i + 1] = h
        return out


class ClipTextTower:
    """CLIPTextModel / CLIPTextModelWithProjection forward ([3P] transformers) on the native kernels."""

    def __init__(self, cfg: ClipTowerConfig, state_dict: Dict[str, torch.Tensor], device="cuda"):
        sd = state_dict
        self.cfg, self.device = cfg, torch.device(device)
        pre = "text_model."
        self.tok = _dev16(sd[pre + "embeddings.token_embedding.weight"], device)
        self.pos = _dev16(sd[pre + "embeddings.position_embedding.weight"], device)
        self.enc = _ClipEncoder(cfg, sd, pre, device)
        self.final_ln = (_dev16(sd[pre + "final_layer_norm.weight"], device), _dev16(sd[pre + "final_layer_norm.bias"], device))
        self.proj = _dev16(sd["text_projection.weight"], device) if "text_projection.weight" in sd else None

    @classmethod
    def from_hf(cls, model, device="cuda") -> "ClipTextTower":
        return cls(ClipTowerConfig.from_hf(model.config), model.state_dict(), device)

    def eos_positions(self, ids: torch.Tensor) -> List[int]:
        """[3P] CLIPTextTransformer pooling: legacy configs (eos_token_id == 2) take argmax(ids) -- the end-of-text
        token has the largest id of the CLIP vocabulary --, newer ones the first position holding eos_token_id."""
        ids = ids.cpu()
        if self.cfg.eos_token_id == 2:
            return ids.argmax(dim=-1).tolist()
        return (ids == self.cfg.eos_token_id).int().argmax(dim=-1).tolist()

    @torch.no_grad()
    def __call__(self, input_ids: torch.Tensor, output_hidden_states: bool = True, hidden_state_index: int = -2):
        """input_ids [B, T] (host or device, any integer dtype).  Returns a namespace with `hidden_states` (dict: the
        requested index -> [B, T, C]; `penultimate` is an alias for hidden_states[-2]), `last_hidden_state`,
        `pooler_output` and, with a projection head, `text_embeds` [B, projection_dim]."""
        B, T = input_ids.shape
        C = self.cfg.hidden_size
        ids32 = input_ids.to(device=self.device, dtype=torch.int32).contiguous()
        h = ops.embed_tokens(ids32, self.tok, self.pos)
        hs = self.enc.run(h, B, T, causal=True, keep=(hidden_state_index,))
        nl = len(self.enc.layers)
        last = ops.layernorm(hs[nl], self.final_ln[0], self.final_ln[1], self.cfg.layer_norm_eps).reshape(B, T, C)
        sel = torch.empty((B, C), dtype=_DTYPE[0], device=self.device)
        for b, pos in enumerate(self.eos_positions(input_ids)):
            sel[b].copy_(last[b, pos])
        out = SimpleNamespace(last_hidden_state=last, pooler_output=sel, text_embeds=None)
        idx = hidden_state_index if hidden_state_index >= 0 else nl + 1 + hidden_state_index
        out.penultimate = hs[idx].reshape(B, T, C)
        out.hidden_states = {hidden_state_index: out.penultimate}
        if self.proj is not None:
            out.text_embeds = ops.linear_small(sel, self.proj)
        return out


class ClipVisionTower:
    """CLIPVisionModelWithProjection forward ([3P] transformers) on the native kernels.  Attribute `config` carries
    `projection_dim` / `hidden_size` like the HF model the reference reads them from (ip_adapter.py:93, :395)."""

    def __init__(self, cfg: ClipTowerConfig, state_dict: Dict[str, torch.Tensor], device="cuda"):
        sd = state_dict
        self.config = cfg
        self.cfg, self.device = cfg, torch.device(device)
        pre = "vision_model."
        C, P = cfg.hidden_size, cfg.patch_size
        self.grid = cfg.image_size // P
        self.kdim = cfg.num_channels * P * P
        self.kpad = (self.kdim + 7) // 8 * 8
        w = sd[pre + "embeddings.patch_embedding.weight"].detach().float().reshape(C, self.kdim)   # [C, c*P*P + py*P + px]
        wp = torch.zeros((C, self.kpad), dtype=torch.float32)
        wp[:, : self.kdim] = w
        self.w_patch = _dev16(wp, device)
        pos = sd[pre + "embeddings.position_embedding.weight"].detach().float()
        cls_tok = sd[pre + "embeddings.class_embedding"].detach().float()
        self.cls_row = _dev16((cls_tok + pos[0]).reshape(1, C), device)          # class token + its position embedding
        self.pos_patches = _dev16(pos[1:], device)                               # fused as the patch GEMM's residual
        self.pre_ln = (_dev16(sd[pre + "pre_layrnorm.weight"], device), _dev16(sd[pre + "pre_layrnorm.bias"], device))
        self.enc = _ClipEncoder(cfg, sd, pre, device)
        self.post_ln = (_dev16(sd[pre + "post_layernorm.weight"], device), _dev16(sd[pre + "post_layernorm.bias"], device))
        self.proj = _dev16(sd["visual_projection.weight"], device) if "visual_projection.weight" in sd else None

    @classmethod
    def from_hf(cls, model, device="cuda") -> "ClipVisionTower":
        return cls(ClipTowerConfig.from_hf(model.config), model.state_dict(), device)

    def to(self, *a, **k):       # IPAdapter.__init__ chains .to(device, dtype=...) on the encoder (ip_adapter.py:81-83)
        return self

    def patch_rows(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """[B, 3, S, S] normalised pixels (host or device) -> [B * grid^2, kpad] fp16 patch rows on the device."""
        B = pixel_values.shape[0]
        P, g = self.cfg.patch_size, self.grid
        x = pixel_values.detach().to("cpu", torch.float32)
        if x.shape[-1] != self.cfg.image_size or x.shape[-2] != self.cfg.image_size:
            raise IHError(f"CLIP vision tower expects {self.cfg.image_size}^2 pixel_values, got {tuple(x.shape)}")
        rows = x.reshape(B, -1, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, self.kdim)
        padded = torch.zeros((B * g * g, self.kpad), dtype=_DTYPE[0])
        padded[:, : self.kdim] = rows.to(_DTYPE[0])
        return padded.to(self.device)

    @torch.no_grad()
    def forward_rows(self, rows: torch.Tensor, B: int, output_hidden_states: bool = False, hidden_state_index: int = -2):
        cfg = self.cfg
        C, g2 = cfg.hidden_size, self.grid * self.grid
        N = g2 + 1
        h = torch.empty((B, N, C), dtype=_DTYPE[0], device=self.device)
        for b in range(B):
            h[b, 0:1].copy_(self.cls_row)
            ops.linear(rows[b * g2:(b + 1) * g2], self.w_patch, residual=self.pos_patches, out=h[b, 1:])
        x = ops.layernorm(h.reshape(B * N, C), self.pre_ln[0], self.pre_ln[1], cfg.layer_norm_eps)
        hs = self.enc.run(x, B, N, causal=False, keep=(hidden_state_index,) if output_hidden_states else ())
        nl = len(self.enc.layers)
        cls_tok = hs[nl].reshape(B, N, C)[:, 0].contiguous()
        pooled = ops.layernorm(cls_tok, self.post_ln[0], self.post_ln[1], cfg.layer_norm_eps)
        out = SimpleNamespace(last_hidden_state=hs[nl].reshape(B, N, C), pooler_output=pooled, image_embeds=None,
                              hidden_states=None)
        if output_hidden_states:
            idx = hidden_state_index if hidden_state_index >= 0 else nl + 1 + hidden_state_index
            out.hidden_states = {hidden_state_index: hs[idx].reshape(B, N, C)}
        if self.proj is not None:
            out.image_embeds = ops.linear_small(pooled, self.proj)
        return out

    def __call__(self, pixel_values: torch.Tensor, output_hidden_states: bool = False):
        return self.forward_rows(self.patch_rows(pixel_values), pixel_values.shape[0], output_hidden_states)


class ClipScorer:
    """PNS judge: score_i = cos(image_embeds(decoded candidate i), text_embeds(prompt)) with the ViT-bigG vision tower
    the adapter already owns and the bigG text tower (SDXL text_encoder_2) -- one joint embedding space, no extra model.
    `decode` maps candidate latents to images in [-1, 1] (the native VAE decoder); the 224^2 CLIP input is produced on
    the device by `ih_resize_patchify_f16` (area average; the PIL bicubic resize of CLIPImageProcessor is a host path)."""

    def __init__(self, vision: ClipVisionTower, text: ClipTextTower, tokenizer=None, decode=None):
        if vision.proj is None or text.proj is None or vision.cfg.projection_dim != text.cfg.projection_dim:
            raise IHError("ClipScorer needs vision and text towers with projection heads into the same space")
        self.vision, self.text, self.tokenizer, self.decode = vision, text, tokenizer, decode
        self._text_embed = None

    def describe(self) -> str:
        c = self.vision.cfg
        return (f"CLIP image-text cosine (native ViT {c.hidden_size}x{c.num_hidden_layers} vision tower + text tower, "
                f"projection {c.projection_dim})")

    @torch.no_grad()
    def set_prompt(self, prompt=None, input_ids: Optional[torch.Tensor] = None) -> None:
        if input_ids is None:
            if self.tokenizer is None:
                raise IHError("ClipScorer.set_prompt needs a tokenizer or input_ids")
            input_ids = self.tokenizer([prompt], padding="max_length", max_length=self.text.cfg.max_position_embeddings,
                                       truncation=True, return_tensors="pt").input_ids
        t = self.text(input_ids).text_embeds.float()
        self._text_embed = t / t.norm(dim=-1, keepdim=True)

    @torch.no_grad()
    def score_images(self, images: torch.Tensor) -> torch.Tensor:
        """images NCHW fp16 in [-1, 1] on the device -> fp32 [B] cosine similarities with the prompt."""
        if self._text_embed is None:
            raise IHError("ClipScorer.set_prompt(...) has not been called")
        v = self.vision
        rows = ops.resize_patchify(images.contiguous(), v.cfg.image_size, v.cfg.patch_size, v.kpad, CLIP_MEAN, CLIP_STD)
        e = v.forward_rows(rows, images.shape[0]).image_embeds.float()
        e = e / e.norm(dim=-1, keepdim=True)
        return (e @ self._text_embed.t())[:, 0]

    def __call__(self, latents: torch.Tensor) -> torch.Tensor:
        if self.decode is None:
            raise IHError("ClipScorer was built without a latent decoder: call score_images(images)")
        return self.score_images(self.decode(latents))
