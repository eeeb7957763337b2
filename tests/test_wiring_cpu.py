"""CPU test of the native UNet / processor / denoise-loop WIRING: imagharmony_b200.ops is swapped for plain-PyTorch fp32
stand-ins (tests/fake_ops.py, test-only) so that layouts, skip-connection order, weight packing, temb offsets and the
K/V cache logic are compared with the oracle to fp32 round-off -- no kernels involved (those are the -m gpu tests)."""
import sys

import pytest
import torch

import fake_ops


@pytest.fixture()
def patched(monkeypatch):
    import imagharmony_b200.ops as real_ops
    for name in dir(fake_ops):
        if not name.startswith("_") and callable(getattr(fake_ops, name)) and hasattr(real_ops, name):
            monkeypatch.setattr(real_ops, name, getattr(fake_ops, name))
    yield


def _build(cfg, seed):
    from imagharmony_b200.unet import UNet2DConditionModel
    from imagharmony_b200.weights import random_state_dict, shapes_of
    from oracle import adapter_ref as A
    from oracle.unet_ref import UNetRef
    with torch.device("meta"):
        shapes = shapes_of(UNetRef(cfg))
    sd = {k: v.float() for k, v in random_state_dict(shapes, seed).items()}
    with torch.device("meta"):
        native = UNet2DConditionModel(cfg)
    native.load_state_dict(sd, assign=True)
    native.install_default_processors()
    procs = torch.nn.ModuleList(native.attn_processors.values()).float()
    ip_sd = {k: v.float() for k, v in random_state_dict(shapes_of(procs), seed + 1).items()}
    procs.load_state_dict(ip_sd)
    native.finalize()
    ref = UNetRef(cfg)
    ref.load_state_dict(sd)
    pr = A.install_processors(ref, cfg)
    torch.nn.ModuleList(pr.values()).load_state_dict(ip_sd)
    return native, ref.eval()


def _inputs(cfg, n, lat, seed=3):
    g = torch.Generator("cpu").manual_seed(seed)
    B = 2 * n
    return (torch.randn(B, 4, lat, lat, generator=g), torch.randn(B, 77 + cfg.num_ip_tokens, cfg.cross_attention_dim, generator=g),
            torch.randn(B, cfg.pooled_embed_dim, generator=g),
            torch.tensor([[lat * 8., lat * 8., 0., 0., lat * 8., lat * 8.]] * B))


def test_native_unet_wiring_matches_oracle(patched):
    from imagharmony_b200.config import TINY
    native, ref = _build(TINY, 0)
    sample, ehs, te, tid = _inputs(TINY, 1, 16)
    with torch.no_grad():
        r = ref(sample, 321.0, ehs, te, tid)
        o = native(sample, torch.full((2,), 321.0), ehs, te, tid)
    assert o.shape == r.shape
    assert torch.allclose(o, r, rtol=1e-4, atol=1e-4), (o - r).abs().max()
    # 140 processors, same names/order as the oracle (and diffusers): the ip_adapter.bin index contract
    from oracle.unet_ref import UNetRef
    assert list(native.attn_processors.keys()) == list(ref.attn_processors.keys())
    assert set(native.state_dict().keys()) == set(ref.state_dict().keys())


def test_native_denoise_loop_wiring_matches_oracle(patched):
    from imagharmony_b200.config import TINY
    from imagharmony_b200.denoise import DenoiseEngine
    from oracle.scheduler_ref import denoise_loop, euler_tables, prepare_latents
    native, ref = _build(TINY, 2)
    T, n, lat = 3, 2, 8
    _, _, ins = euler_tables(T)
    latents = prepare_latents(n, 4, lat, lat, [1, 2], ins, dtype=torch.float32)
    _, ehs, te, tid = _inputs(TINY, n, lat, seed=5)
    procs = [p for p in ref.attn_processors.values() if hasattr(p, "to_k_ip")]

    def set_scale(s):
        for p in procs:
            p.scale = s
    r = denoise_loop(lambda s, t, e, x, y: ref(s, t, e, x, y), latents.clone(), ehs[n:], ehs[:n], te[n:], te[:n], tid[:n], T,
                     guidance_scale=5.0, set_scale=set_scale, conditioning_scale=0.6, control_guidance_start=0.3)
    eng = DenoiseEngine.__new__(DenoiseEngine)
    DenoiseEngine.__init__(eng, native, use_cuda_graph=False)
    # fp32 CPU buffers for the wiring test
    import imagharmony_b200.denoise as dn
    orig_empty = torch.empty

    def empty32(*a, **k):
        if k.get("dtype") == torch.float16:
            k["dtype"] = torch.float32
        return orig_empty(*a, **k)
    dn.torch.empty = empty32
    try:
        o = eng.run(latents.clone(), ehs[n:], ehs[:n], te[n:], te[:n], tid[:n], T, guidance_scale=5.0, ip_scale=0.6,
                    control_guidance_start=0.3)
    finally:
        dn.torch.empty = orig_empty
    assert torch.allclose(o, r, rtol=2e-4, atol=2e-4), (o - r).abs().max()
