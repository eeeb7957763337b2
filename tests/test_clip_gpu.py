"""Scope row f2 on the GPU: the native CLIP towers / PNS judge (imagharmony_b200/clip.py on the sm_100a kernels through the
C ABI) against the `transformers` classes the reference uses (ip_adapter.py:81-84,163-164; encode_prompt :292-319) in fp32
on the CPU, at miniature AND full-size configurations (CLIP ViT-L text, OpenCLIP bigG text, ViT-bigG/14 vision), random
init with fp16-representable weights.  Deep-stack bar as in test_unet_gpu.py: the native error must not exceed twice the
error of the same transformers model run in fp16 on the GPU (floor 2e-3 of the output range)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
pytest.importorskip("transformers")


def _close(name, got, want, eager=None, tol=2e-3):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs().max().item()
    mx = want.abs().max().item()
    e_eag = (eager.float().cpu() - want).abs().max().item() if eager is not None else 0.0
    print(f"[{name}] native max|err| {err:.3e}  hf-fp16 {e_eag:.3e}  max|ref| {mx:.3e}")
    assert torch.isfinite(got).all() and err <= max(2.0 * e_eag, tol * mx), (name, err, e_eag, mx)


@pytest.mark.parametrize("B,H,Nq,Nk,d,causal", [(2, 12, 77, 77, 64, True), (1, 16, 257, 257, 104, False),
                                                 (2, 20, 77, 77, 64, True), (3, 4, 19, 33, 40, False), (1, 2, 9, 12, 16, True)])
def test_attention_generic(B, H, Nq, Nk, d, causal):
    from imagharmony_b200 import ops
    g = torch.Generator().manual_seed(B + H + Nq)
    q = (torch.randn(B * Nq, 3 * H * d, generator=g) * 1.5).half()
    kv = (torch.randn(B * Nk, 3 * H * d, generator=g) * 1.5).half()
    qh = q[:, :H * d].float().reshape(B, Nq, H, d).permute(0, 2, 1, 3)
    kh = kv[:, H * d:2 * H * d].float().reshape(B, Nk, H, d).permute(0, 2, 1, 3)
    vh = kv[:, 2 * H * d:].float().reshape(B, Nk, H, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) * d ** -0.5
    if causal:
        s = s.masked_fill(~torch.ones(Nq, Nk, dtype=torch.bool).tril(Nk - Nq), float("-inf"))
    want = (s.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B * Nq, H * d)
    qd, kvd = q.cuda(), kv.cuda()
    out = ops.attention_generic(qd[:, :H * d], kvd[:, H * d:2 * H * d], kvd[:, 2 * H * d:], B, H, Nq, Nk, d, d, d ** -0.5, causal)
    torch.cuda.synchronize()
    _close(f"attention_generic hd{d} causal={causal}", out, want, tol=1e-3)


def test_embed_quickgelu_resize_kernels():
    from imagharmony_b200 import ops
    from imagharmony_b200.clip import CLIP_MEAN, CLIP_STD
    from oracle.clip_ref import resize_patchify_ref
    g = torch.Generator().manual_seed(0)
    tok, pos = torch.randn(300, 64, generator=g).half(), torch.randn(77, 64, generator=g).half()
    ids = torch.randint(0, 300, (3, 77), generator=g)
    out = ops.embed_tokens(ids.int().cuda(), tok.cuda(), pos.cuda())
    _close("embed_tokens", out, (tok.float()[ids] + pos.float()[None]).reshape(3 * 77, 64), tol=1e-3)
    x, w, b = torch.randn(200, 96, generator=g).half(), (torch.randn(160, 96, generator=g) * 0.1).half(), torch.randn(160, generator=g).half()
    y = x.float() @ w.float().t() + b.float()
    _close("gemm quick_gelu epilogue", ops.linear(x.cuda(), w.cuda(), b.cuda(), quick_gelu=True), y * torch.sigmoid(1.702 * y), tol=1e-3)
    for shape, size in (((2, 3, 1024, 1024), 224), ((1, 3, 100, 132), 56), ((1, 3, 224, 224), 224)):
        img = (torch.rand(shape, generator=g) * 2.2 - 1.1).half()
        rows = ops.resize_patchify(img.cuda(), size, 14, 592, CLIP_MEAN, CLIP_STD)
        _close(f"resize_patchify {shape}->{size}", rows, resize_patchify_ref(img.float(), size, 14, 592, CLIP_MEAN, CLIP_STD), tol=2e-3)


def _ids(B, vocab, eos, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 2, (B, 77), generator=g)
    ids[:, 0] = vocab - 2
    for b in range(B):
        n = 6 + 9 * b
        ids[b, n:] = eos
    return ids


@pytest.mark.parametrize("name,with_proj", [("mini", True), ("TEXT_L", False), ("TEXT_BIGG", True)])
def test_text_tower_matches_transformers(name, with_proj):
    """SDXL text_encoder (CLIP ViT-L: 12 x 768, quick_gelu, hidden_states[-2]) and text_encoder_2 (OpenCLIP bigG: 32 x 1280,
    gelu, projected text_embeds) at full size, plus a miniature."""
    from imagharmony_b200.clip import ClipTextTower
    from oracle import clip_ref as R
    kw = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, hidden_act="quick_gelu",
              projection_dim=64) if name == "mini" else getattr(R, name)
    hf = R.hf_text_model(11, with_proj, **kw)
    tower = ClipTextTower.from_hf(hf, device="cuda")
    ids = _ids(2, 49408, 49407, 3)
    with torch.no_grad():
        want = hf(ids, output_hidden_states=True)
        eag = hf.half().cuda()(ids.cuda(), output_hidden_states=True)
    got = tower(ids)
    torch.cuda.synchronize()
    _close(f"{name} hidden_states[-2]", got.penultimate, want.hidden_states[-2], eag.hidden_states[-2])
    if with_proj:
        _close(f"{name} text_embeds", got.text_embeds, want.text_embeds, eag.text_embeds)
    else:
        _close(f"{name} pooler_output", got.pooler_output, want.pooler_output, eag.pooler_output)


@pytest.mark.parametrize("name", ["mini", "VISION_BIGG"])
def test_vision_tower_matches_transformers(name):
    """The IP-Adapter SDXL image encoder: OpenCLIP ViT-bigG/14 (48 x 1664, 16 heads of 104, 257 tokens) -> image_embeds
    [1, 1280] (ip_adapter.py:163-164) and hidden_states[-2] [1, 257, 1664] (Plus variant, :404-412)."""
    from imagharmony_b200.clip import ClipVisionTower
    from oracle import clip_ref as R
    kw = dict(hidden_size=208, intermediate_size=320, num_hidden_layers=3, num_attention_heads=2, hidden_act="gelu",
              projection_dim=40, image_size=56, patch_size=14) if name == "mini" else R.VISION_BIGG
    hf = R.hf_vision_model(13, **kw)
    tower = ClipVisionTower.from_hf(hf, device="cuda")
    S = kw["image_size"]
    px = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(5)).half().float()
    with torch.no_grad():
        want = hf(px, output_hidden_states=True)
        eag = hf.half().cuda()(px.half().cuda(), output_hidden_states=True)
    got = tower(px, output_hidden_states=True)
    torch.cuda.synchronize()
    _close(f"{name} image_embeds", got.image_embeds, want.image_embeds, eag.image_embeds)
    _close(f"{name} hidden_states[-2]", got.hidden_states[-2], want.hidden_states[-2], eag.hidden_states[-2])


def test_clip_scorer_and_pns_on_gpu():
    """The PNS judge end to end: decoded images -> device-side resize/normalise/patchify -> vision tower -> cosine with the
    prompt's text embedding; checked against transformers fp32 on the same (oracle-)preprocessed pixels, and used by
    pns_select to pick a winner."""
    from imagharmony_b200.clip import CLIP_MEAN, CLIP_STD, ClipScorer, ClipTextTower, ClipVisionTower
    from imagharmony_b200.pns import pns_select
    from oracle import clip_ref as R
    hv = R.hf_vision_model(6, hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4,
                           hidden_act="gelu", projection_dim=64, image_size=224, patch_size=14)
    ht = R.hf_text_model(7, True, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                         hidden_act="gelu", projection_dim=64)
    scorer = ClipScorer(ClipVisionTower.from_hf(hv, device="cuda"), ClipTextTower.from_hf(ht, device="cuda"),
                        decode=lambda lat: F.interpolate(lat[:, :3].float(), scale_factor=8).clamp(-1, 1).half())
    ids = _ids(1, 49408, 49407, 9)
    scorer.set_prompt(input_ids=ids)
    lat = torch.randn(4, 4, 32, 32, generator=torch.Generator().manual_seed(2)).half().cuda()
    imgs = scorer.decode(lat)
    got = scorer.score_images(imgs)
    rows = R.resize_patchify_ref(imgs.float().cpu(), 224, 14, 592, CLIP_MEAN, CLIP_STD)[:, :588]
    px = rows.reshape(4, 16, 16, 3, 14, 14).permute(0, 3, 1, 4, 2, 5).reshape(4, 3, 224, 224)
    with torch.no_grad():
        want = F.cosine_similarity(hv(px).image_embeds, ht(ids).text_embeds.expand(4, -1))
    print(f"[ClipScorer] native {got.tolist()} vs transformers {want.tolist()}")
    assert torch.allclose(got.cpu(), want, atol=5e-3)
    res = pns_select(lambda seeds: lat, [10, 11, 12, 13], scorer, max_batch=4)
    assert res.best_index == int(want.argmax()) and torch.equal(res.best_latents, lat[res.best_index])
