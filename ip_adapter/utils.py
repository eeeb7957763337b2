"""Host-side helpers with the reference's names (ip_adapter/utils.py).

* `is_torch2_available`, `get_generator` (:80-93): used by the hot path (seeds -> generators).
* The attention-map diagnostics (:6-79): `register_cross_attention_hook`, `get_net_attn_map`, `attnmaps2images`, `upscale`
  and the module-level `attn_maps` store.  In the reference every non-skip `IPAttnProcessor2_0` call computes `attn_map`
  (attention_processor.py:443-444) whether or not anybody collects it; here the processors compute it only after
  `register_cross_attention_hook(unet)` switched them on (`keep_attn_map`), so the timed path never pays for it.  With the
  CUDA-graphed loop the hooks run at capture time and the stored tensors are the graph's own buffers: after a `generate`
  they hold the maps of the last denoise step, which is also what the reference's per-step overwrite leaves behind.
"""
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

attn_maps = {}          # attn2 module name -> [B, heads, N, num_tokens] map of its last call (filled by the hooks)


def hook_fn(name):
    """Forward hook of one attn2 module: move the processor's `attn_map` into `attn_maps[name]` (utils.py:7-13)."""
    def forward_hook(module, args, output):
        proc = module.processor
        if hasattr(proc, "attn_map"):
            attn_maps[name] = proc.attn_map
            del proc.attn_map
    return forward_hook


def register_cross_attention_hook(unet):
    """Hook every cross-attention (`...attn2`) module of `unet` (utils.py:15-20) and switch the map on in the processors
    that can produce it."""
    for name, module in unet.named_modules():
        if name.rsplit(".", 1)[-1].startswith("attn2"):
            module.register_forward_hook(hook_fn(name))
            if hasattr(getattr(module, "processor", None), "keep_attn_map"):
                module.processor.keep_attn_map = True
    if hasattr(unet, "graph_epoch"):
        unet.graph_epoch += 1      # step graphs captured without the maps are stale (DenoiseEngine re-captures)
    return unet


def upscale(attn_map, target_size):
    """[heads, N, tokens] -> [tokens, H, W]: head mean, the token axis first, the N latent positions back on their grid
    (the level is found from N: the map lives on target/8/2^i), bilinear resize to `target_size`, softmax over the tokens
    (utils.py:22-45)."""
    per_token = attn_map.mean(dim=0).transpose(0, 1)                  # [tokens, N]
    n = per_token.shape[1]
    grid = None
    for level in range(5):
        s = 2 ** level
        if (target_size[0] // s) * (target_size[1] // s) == n * 64:
            grid = (target_size[0] // (8 * s), target_size[1] // (8 * s))
            break
    assert grid is not None, "temp_size cannot is None"
    planes = per_token.reshape(per_token.shape[0], *grid)
    planes = F.interpolate(planes[None].float(), size=target_size, mode="bilinear", align_corners=False)[0]
    return planes.softmax(dim=0)


def get_net_attn_map(image_size, batch_size=2, instance_or_negative=False, detach=True):
    """Mean over the hooked layers of the upscaled maps of ONE CFG half (chunk 1 = conditional unless
    `instance_or_negative`) (utils.py:46-59)."""
    half = 0 if instance_or_negative else 1
    layers = []
    for stored in attn_maps.values():
        m = stored.cpu() if detach else stored
        layers.append(upscale(torch.chunk(m, batch_size)[half].squeeze(), image_size))
    return torch.stack(layers, dim=0).mean(dim=0)


def attnmaps2images(net_attn_maps):
    """One 8-bit grey image per token, min-max normalised (utils.py:61-79)."""
    images = []
    for plane in net_attn_maps:
        a = plane.cpu().numpy()
        a = (a - a.min()) / (a.max() - a.min()) * 255
        images.append(Image.fromarray(a.astype(np.uint8)))
    return images


def is_torch2_available():
    return hasattr(F, "scaled_dot_product_attention")


def get_generator(seed, device):
    """seed -> torch.Generator; a list of seeds -> a list of generators, one per image (utils.py:83-93).  One generator
    per image is what makes a PNS candidate noise identical on whichever rank / batch slot it lands."""
    def make(value):
        return torch.Generator(device).manual_seed(value)
    if seed is None:
        return None
    return [make(v) for v in seed] if isinstance(seed, list) else make(seed)
