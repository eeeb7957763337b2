"""CPU restatement of the decoder half of [3P] diffusers==0.30.0 `AutoencoderKL` (requirements.txt:25; not vendored) as
the reference calls it: custom_pipelines.py:373 `self.vae.decode(latents / self.vae.config.scaling_factor)` followed by
`image_processor.postprocess` (:383).  TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu arm).  **Parity unpinned**:
diffusers is not installable here and the reference holds no golden vector for the VAE; the structure follows the
public SDXL `vae/config.json` (block_out_channels 128/256/512/512, layers_per_block 2, 32 groups, eps 1e-6, one
single-head attention of width 512 in the mid block, nearest 2x upsampling, scaling_factor 0.13025).
State-dict keys equal diffusers' (`post_quant_conv.*`, `decoder.*`)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetRef(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))            # dropout p = 0, no time embedding in the VAE
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h                                     # output_scale_factor = 1


class AttentionRef(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        res = x
        h = self.group_norm(x).reshape(B, C, H * W).transpose(1, 2)          # [B, N, C]
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)                   # heads = 1, dim_head = C
        a = torch.softmax(q @ k.transpose(1, 2) * (C ** -0.5), dim=-1)
        o = self.to_out[0](a @ v)
        return o.transpose(1, 2).reshape(B, C, H, W) + res                   # residual_connection, rescale 1


class UpsampleRef(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class MidRef(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.attentions = nn.ModuleList([AttentionRef(ch, groups)])
        self.resnets = nn.ModuleList([ResnetRef(ch, ch, groups), ResnetRef(ch, ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class UpBlockRef(nn.Module):
    def __init__(self, cin, cout, n_res, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetRef(cin if j == 0 else cout, cout, groups) for j in range(n_res)])
        if add_up:
            self.upsamplers = nn.ModuleList([UpsampleRef(cout)])
        self.has_up = add_up

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if self.has_up else x


class DecoderRef(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch = cfg.decoder_channels
        g = cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[0], 3, padding=1)
        self.mid_block = MidRef(ch[0], g)
        ups, prev = [], ch[0]
        for i, c in enumerate(ch):
            ups.append(UpBlockRef(prev, c, cfg.layers_per_block + 1, g, add_up=i < len(ch) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class VAEDecoderRef(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = DecoderRef(cfg)

    def decode(self, latents):
        """custom_pipelines.py:373: vae.decode(latents / scaling_factor)[0]"""
        return self.decoder(self.post_quant_conv(latents / self.config.scaling_factor))


def postprocess_ref(image: torch.Tensor):
    """VaeImageProcessor.postprocess(output_type='np'): (x / 2 + 0.5).clamp(0, 1) -> NHWC float numpy"""
    return (image.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).cpu().numpy()


def tiled_decode_ref(vae: "VAEDecoderRef", latents: torch.Tensor) -> torch.Tensor:
    """[3P] AutoencoderKL.tiled_decode (test.py:73 `pipe.enable_vae_tiling()`): 25 % overlapping tiles of
    tile_latent_min_size latents, each through post_quant_conv + decoder on its own, blended with the upper and the left
    neighbour over `blend_extent` pixels and cropped to `row_limit`."""
    cfg = vae.config
    z = latents / cfg.scaling_factor
    tl = cfg.tile_latent_min_size
    overlap_size = int(tl * (1 - cfg.tile_overlap_factor))
    blend_extent = int(cfg.sample_size * cfg.tile_overlap_factor)
    row_limit = cfg.sample_size - blend_extent
    rows = []
    for i in range(0, z.shape[2], overlap_size):
        row = []
        for j in range(0, z.shape[3], overlap_size):
            row.append(vae.decoder(vae.post_quant_conv(z[:, :, i:i + tl, j:j + tl])))
        rows.append(row)
    result_rows = []
    for i, row in enumerate(rows):
        result_row = []
        for j, tile in enumerate(row):
            if i > 0:
                a, b = rows[i - 1][j], tile
                ext = min(a.shape[2], b.shape[2], blend_extent)
                for y in range(ext):
                    b[:, :, y, :] = a[:, :, -ext + y, :] * (1 - y / ext) + b[:, :, y, :] * (y / ext)
            if j > 0:
                a, b = row[j - 1], tile
                ext = min(a.shape[3], b.shape[3], blend_extent)
                for x in range(ext):
                    b[:, :, :, x] = a[:, :, :, -ext + x] * (1 - x / ext) + b[:, :, :, x] * (x / ext)
            result_row.append(tile[:, :, :row_limit, :row_limit])
        result_rows.append(torch.cat(result_row, dim=3))
    return torch.cat(result_rows, dim=2)
