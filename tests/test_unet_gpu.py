"""Model-level GPU parity: native processors / UNet / denoise loop (sm_100a kernels through the C ABI) against the
CPU fp32 oracle (oracle/) on identical seeded inputs and fp16-representable weights, and against the golden vectors
generated from the reference's own modules (tests/golden/reference_vectors.pt).

Tolerances.  Single ops: rtol = atol = 1e-3 (BASELINE.json).  Whole-UNet / trajectory outputs pass through ~100
fp16 roundings, so the bar there is the one SURVEY.md section 7 defines: the native path's error against the fp32 oracle
must not exceed the error of the *same oracle run in torch-eager fp16 on the GPU* (the stand-in for the reference's GPU
diffusers path) by more than a factor 2, with an absolute floor of 2e-3 * max|ref|.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.pt")


def _make_pair(cfg, seed=0, ip_seed=1):
    """(native unet on cuda fp16, oracle unet on cpu fp32, oracle unet on cuda fp16) with identical parameters."""
    from imagharmony_b200.unet import UNet2DConditionModel
    from imagharmony_b200.weights import random_state_dict, shapes_of
    from oracle import adapter_ref as A
    from oracle.unet_ref import UNetRef

    with torch.device("meta"):
        shapes = shapes_of(UNetRef(cfg))
    sd = random_state_dict(shapes, seed)
    native = UNet2DConditionModel.from_state_dict(cfg, sd, device="cuda")
    procs = torch.nn.ModuleList(native.attn_processors.values())
    ip_sd = random_state_dict(shapes_of(procs), ip_seed)
    procs.load_state_dict({k: v.cuda() for k, v in ip_sd.items()})
    native.finalize()

    def oracle(device, dtype):
        m = UNetRef(cfg)
        m.load_state_dict({k: v.float() for k, v in sd.items()})
        pr = A.install_processors(m, cfg, dtype=torch.float32)
        torch.nn.ModuleList(pr.values()).load_state_dict({k: v.float() for k, v in ip_sd.items()})
        return m.to(device=device, dtype=dtype).eval()

    return native, oracle("cpu", torch.float32), oracle("cuda", torch.float16)


def _inputs(cfg, n_img, res_lat, seed=3):
    g = torch.Generator("cpu").manual_seed(seed)
    B = 2 * n_img
    L = 77 + cfg.num_ip_tokens
    return {
        "sample": torch.randn(B, 4, res_lat, res_lat, generator=g).half(),
        "ehs": torch.randn(B, L, cfg.cross_attention_dim, generator=g).half(),
        "text_embeds": torch.randn(B, cfg.pooled_embed_dim, generator=g).half(),
        "time_ids": torch.tensor([[res_lat * 8., res_lat * 8., 0., 0., res_lat * 8., res_lat * 8.]] * B),
    }


def _errs(native_out, ref32, eager16):
    ref = ref32.float().cpu()
    e_nat = (native_out.float().cpu() - ref).abs().max().item()
    e_eag = (eager16.float().cpu() - ref).abs().max().item()
    return e_nat, e_eag, ref.abs().max().item()


def test_native_processors_match_reference_goldens():
    """IPAttnProcessor2_0 / AttnProcessor2_0 on the kernels vs outputs of the reference's own classes."""
    from imagharmony_b200.unet import Attention
    from ip_adapter.attention_processor import AttnProcessor2_0, IPAttnProcessor2_0
    gold = torch.load(GOLDEN, map_location="cpu")
    for skip in (0, 1):
        g = gold[f"ipattn_skip{skip}"]
        C, D = g["hidden"].shape[-1], g["ehs"].shape[-1]
        attn = Attention(C, g["heads"], D)
        attn.load_state_dict(g["attn"])
        attn = attn.half().cuda()
        proc = IPAttnProcessor2_0(C, D, scale=g["scale"], num_tokens=g["num_tokens"], skip=bool(skip))
        proc.load_state_dict(g["proc"])
        proc = proc.half().cuda()
        # the golden inputs are fp32; compare against the reference evaluated on the same fp16-rounded operands
        from oracle import adapter_ref as A
        from oracle.unet_ref import Attention as RefAttention
        ra = RefAttention(C, g["heads"], D)
        ra.load_state_dict({k: v.half().float() for k, v in g["attn"].items()})
        rp = A.IPAttnProcessorRef(C, D, scale=g["scale"], num_tokens=g["num_tokens"], skip=bool(skip))
        rp.load_state_dict({k: v.half().float() for k, v in g["proc"].items()})
        hid, ehs = g["hidden"].half(), g["ehs"].half()
        ref = rp(ra, hid.float(), encoder_hidden_states=ehs.float())
        out = proc(attn, hid.cuda(), encoder_hidden_states=ehs.cuda())
        torch.cuda.synchronize()
        assert torch.allclose(out.float().cpu(), ref, rtol=2e-3, atol=2e-3), (out.float().cpu() - ref).abs().max()
        # and stays close to the un-rounded golden output of the real reference class
        assert (out.float().cpu() - g["out"]).abs().max() < 2e-2
    g = gold["selfattn"]
    C = g["hidden"].shape[-1]
    attn = Attention(C, g["heads"])
    attn.load_state_dict(g["attn"])
    attn = attn.half().cuda()
    out = AttnProcessor2_0()(attn, g["hidden"].half().cuda())
    assert (out.float().cpu() - g["out"]).abs().max() < 2e-2


def test_unet_forward_tiny_matches_oracle():
    from imagharmony_b200.config import TINY
    native, ref32, eager16 = _make_pair(TINY)
    x = _inputs(TINY, 1, 32)
    t = 500.0
    with torch.no_grad():
        r = ref32(x["sample"].float(), t, x["ehs"].float(), x["text_embeds"].float(), x["time_ids"])
        e = eager16(x["sample"].cuda(), t, x["ehs"].cuda(), x["text_embeds"].cuda(), x["time_ids"].cuda())
        tt = torch.full((2,), t, device="cuda")
        o = native(x["sample"].cuda(), tt, x["ehs"].cuda(), x["text_embeds"].cuda(), x["time_ids"].cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(o).all()
    e_nat, e_eag, mx = _errs(o, r, e)
    print(f"[unet tiny fwd] native err {e_nat:.3e}  eager-fp16 err {e_eag:.3e}  max|ref| {mx:.3e}")
    assert e_nat <= max(2.0 * e_eag, 2e-3 * mx), (e_nat, e_eag, mx)


def test_unet_forward_tiny_odd_sizes_and_batch():
    """768^2-like sizes (latent 24 -> 12 -> 6 tokens per side) and n = 2 images (UNet batch 4)."""
    from imagharmony_b200.config import TINY
    native, ref32, eager16 = _make_pair(TINY, seed=5)
    x = _inputs(TINY, 2, 24)
    with torch.no_grad():
        r = ref32(x["sample"].float(), 37.0, x["ehs"].float(), x["text_embeds"].float(), x["time_ids"])
        e = eager16(x["sample"].cuda(), 37.0, x["ehs"].cuda(), x["text_embeds"].cuda(), x["time_ids"].cuda())
        o = native(x["sample"].cuda(), torch.full((4,), 37.0, device="cuda"), x["ehs"].cuda(), x["text_embeds"].cuda(),
                   x["time_ids"].cuda())
    e_nat, e_eag, mx = _errs(o, r, e)
    print(f"[unet tiny 24x24 n2] native err {e_nat:.3e}  eager-fp16 err {e_eag:.3e}  max|ref| {mx:.3e}")
    assert e_nat <= max(2.0 * e_eag, 2e-3 * mx), (e_nat, e_eag, mx)


@pytest.mark.parametrize("use_graph", [False, True])
def test_denoise_loop_tiny_matches_oracle(use_graph):
    """4-step CFG + Euler trajectory (custom_pipelines.py:325-363), with the IP scale gated off for the last step."""
    from imagharmony_b200.config import TINY
    from imagharmony_b200.denoise import DenoiseEngine
    from oracle.scheduler_ref import denoise_loop, euler_tables, prepare_latents
    native, ref32, eager16 = _make_pair(TINY, seed=7)
    T, n, lat = 4, 1, 32
    _, _, ins = euler_tables(T)
    latents = prepare_latents(n, 4, lat, lat, [42], ins)
    x = _inputs(TINY, n, lat, seed=9)
    neg, pos = x["ehs"][:n], x["ehs"][n:]
    npool, ppool = x["text_embeds"][:n], x["text_embeds"][n:]
    tid = x["time_ids"][:n]

    def run_oracle(m, dev, dt):
        procs = [p for p in m.attn_processors.values() if hasattr(p, "to_k_ip")]

        def set_scale(s):
            for p in procs:
                p.scale = s
        fn = lambda s, t, e, te, ti: m(s.to(dt), t, e, te, ti).to(torch.float16)  # noqa: E731
        return denoise_loop(fn, latents.to(dev), pos.to(dev, dt), neg.to(dev, dt), ppool.to(dev, dt), npool.to(dev, dt),
                            tid.to(dev), T, guidance_scale=5.0, set_scale=set_scale, conditioning_scale=0.8,
                            control_guidance_end=0.75)

    r = run_oracle(ref32, "cpu", torch.float32)
    e = run_oracle(eager16, "cuda", torch.float16)
    eng = DenoiseEngine(native, use_cuda_graph=use_graph)
    o = eng.run(latents.pin_memory(), pos, neg, ppool, npool, tid, T, guidance_scale=5.0, ip_scale=0.8,
                control_guidance_end=0.75)
    torch.cuda.synchronize()
    e_nat, e_eag, mx = _errs(o, r, e)
    print(f"[denoise tiny graph={use_graph}] native err {e_nat:.3e}  eager-fp16 err {e_eag:.3e}  max|ref| {mx:.3e}")
    assert torch.isfinite(o).all()
    assert e_nat <= max(2.0 * e_eag, 2e-3 * mx), (e_nat, e_eag, mx)
    if use_graph:
        # replaying the captured graphs on fresh inputs must give the same answer as the first run
        o2 = eng.run(latents.pin_memory(), pos, neg, ppool, npool, tid, T, guidance_scale=5.0, ip_scale=0.8,
                     control_guidance_end=0.75)
        assert torch.equal(o, o2)
        assert eng.last_launches_per_step > 50
