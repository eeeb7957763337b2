"""Scope row f2 on the CPU: the WIRING of the native CLIP towers (imagharmony_b200/clip.py -- weight packing, q|k|v fusion,
causal mask, EOS pooling, penultimate hidden states, patch GEMM + fused position embedding, projection heads) against the
`transformers` classes the reference uses (ip_adapter.py:81-84,163-164; encode_prompt :292-319), with imagharmony_b200.ops
swapped for the plain-PyTorch fp32 stand-ins of tests/fake_ops.py.  The kernels themselves are the -m gpu tests."""
import pytest
import torch

import fake_ops

pytest.importorskip("transformers")


@pytest.fixture()
def patched(monkeypatch):
    import imagharmony_b200.clip as clip
    import imagharmony_b200.ops as real_ops
    for name in dir(fake_ops):
        if not name.startswith("_") and callable(getattr(fake_ops, name)) and hasattr(real_ops, name):
            monkeypatch.setattr(real_ops, name, getattr(fake_ops, name))
    monkeypatch.setattr(clip, "_DTYPE", [torch.float32])
    yield


def _ids(B, vocab, eos, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 2, (B, 77), generator=g)
    ids[:, 0] = vocab - 2
    for b in range(B):
        n = 5 + 7 * b
        ids[b, n] = eos
        ids[b, n + 1:] = eos            # padded with the end-of-text token like CLIPTokenizer(padding="max_length")
    return ids


@pytest.mark.parametrize("act,proj,eos", [("quick_gelu", False, 2), ("gelu", True, 99), ("gelu", True, 2)])
def test_text_tower_wiring_matches_transformers(patched, act, proj, eos):
    from imagharmony_b200.clip import ClipTextTower
    from oracle.clip_ref import hf_text_model
    vocab = 100
    hf = hf_text_model(3, proj, vocab_size=vocab, eos_token_id=eos, hidden_size=64, intermediate_size=160,
                       num_hidden_layers=3, num_attention_heads=4, hidden_act=act, projection_dim=48)
    tower = ClipTextTower.from_hf(hf, device="cpu")
    ids = _ids(3, vocab, vocab - 1 if eos == 2 else eos, 5)      # legacy (eos id 2) pools at argmax(ids) = the largest id
    with torch.no_grad():
        want = hf(ids, output_hidden_states=True)
    got = tower(ids)
    assert torch.allclose(got.penultimate, want.hidden_states[-2], atol=2e-5), (got.penultimate - want.hidden_states[-2]).abs().max()
    assert torch.allclose(got.last_hidden_state, want.last_hidden_state, atol=2e-5)
    if proj:
        assert torch.allclose(got.text_embeds, want.text_embeds, atol=2e-5), (got.text_embeds - want.text_embeds).abs().max()
    else:
        assert torch.allclose(got.pooler_output, want.pooler_output, atol=2e-5)


def test_vision_tower_wiring_matches_transformers(patched):
    """head_dim 104 like ViT-bigG (hidden 208 / 2 heads), 56^2 image / patch 14 -> 16 patches + class token."""
    from imagharmony_b200.clip import ClipVisionTower
    from oracle.clip_ref import hf_vision_model
    hf = hf_vision_model(4, hidden_size=208, intermediate_size=320, num_hidden_layers=3, num_attention_heads=2,
                         hidden_act="gelu", projection_dim=40, image_size=56, patch_size=14)
    tower = ClipVisionTower.from_hf(hf, device="cpu")
    assert tower.config.projection_dim == 40 and tower.config.hidden_size == 208 and tower.kpad == 592
    px = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = hf(px, output_hidden_states=True)
    got = tower(px, output_hidden_states=True)
    assert torch.allclose(got.image_embeds, want.image_embeds, atol=3e-5), (got.image_embeds - want.image_embeds).abs().max()
    assert torch.allclose(got.hidden_states[-2], want.hidden_states[-2], atol=3e-5)
    assert got.hidden_states[-2].shape == (2, 17, 208)


def test_clip_scorer_wiring(patched):
    """PNS judge = cosine(image_embeds(resized decoded image), text_embeds(prompt)) in the towers' joint space."""
    from imagharmony_b200.clip import CLIP_MEAN, CLIP_STD, ClipScorer, ClipTextTower, ClipVisionTower
    from oracle.clip_ref import hf_text_model, hf_vision_model, resize_patchify_ref
    hv = hf_vision_model(6, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         hidden_act="gelu", projection_dim=32, image_size=56, patch_size=14)
    ht = hf_text_model(7, True, vocab_size=100, eos_token_id=99, hidden_size=48, intermediate_size=96,
                       num_hidden_layers=2, num_attention_heads=2, hidden_act="gelu", projection_dim=32)
    scorer = ClipScorer(ClipVisionTower.from_hf(hv, device="cpu"), ClipTextTower.from_hf(ht, device="cpu"))
    ids = _ids(1, 100, 99, 9)
    scorer.set_prompt(input_ids=ids)
    imgs = torch.rand(3, 3, 100, 132, generator=torch.Generator().manual_seed(2)) * 2 - 1       # non-integer resize ratios
    got = scorer.score_images(imgs)
    rows = resize_patchify_ref(imgs, 56, 14, 592, CLIP_MEAN, CLIP_STD)[:, :588]
    px = rows.reshape(3, 4, 4, 3, 14, 14).permute(0, 3, 1, 4, 2, 5).reshape(3, 3, 56, 56)
    with torch.no_grad():
        e = hv(px).image_embeds
        t = ht(ids).text_embeds
    want = torch.nn.functional.cosine_similarity(e, t.expand_as(e))
    assert torch.allclose(got, want, atol=1e-4), (got, want)
    assert "cosine" in scorer.describe()


def test_resize_patchify_ref_is_identity_resize_at_equal_size():
    from oracle.clip_ref import resize_patchify_ref
    img = torch.rand(1, 3, 28, 28) * 2 - 1
    rows = resize_patchify_ref(img, 28, 14, 592, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))       # (v - .5) / .5 undoes [-1,1] -> [0,1]
    back = rows[:, :588].reshape(1, 2, 2, 3, 14, 14).permute(0, 3, 1, 4, 2, 5).reshape(1, 3, 28, 28)
    assert torch.allclose(back, img, atol=1e-6) and torch.count_nonzero(rows[:, 588:]) == 0


def test_tower_param_shapes_match_transformers_keys():
    """The key / shape table used to build random-init towers on the GPU (bench.py's CLIP judge) is the transformers one."""
    from imagharmony_b200.clip import ClipTowerConfig, tower_param_shapes
    from oracle.clip_ref import hf_text_model, hf_vision_model
    ht = hf_text_model(1, True, vocab_size=100, eos_token_id=99, hidden_size=64, intermediate_size=96,
                       num_hidden_layers=2, num_attention_heads=2, hidden_act="gelu", projection_dim=32)
    hv = hf_vision_model(1, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=2,
                         hidden_act="gelu", projection_dim=32, image_size=56, patch_size=14)
    for hf, kind in ((ht, "text"), (hv, "vision")):
        want = {k: tuple(v.shape) for k, v in hf.state_dict().items() if "position_ids" not in k}
        assert tower_param_shapes(ClipTowerConfig.from_hf(hf.config), kind) == want
