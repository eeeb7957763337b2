"""Prompt conditioning (scope row f2): ClipPromptEncoder restates diffusers' SDXL encode_prompt on top of the NATIVE CLIP text
towers (imagharmony_b200.clip), checked against the transformers CLIP classes.  No weights / vocabularies exist offline, so the test builds miniature random CLIP text models and a toy
character-level CLIP vocabulary."""
import json

import pytest
import torch


def _toy_tokenizer(tmp_path, name):
    from transformers import CLIPTokenizer
    chars = list("abcdefghijklmnopqrstuvwxyz ,")
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
        vocab[c + "</w>"] = len(vocab)
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    d = tmp_path / name
    d.mkdir()
    (d / "vocab.json").write_text(json.dumps(vocab))
    (d / "merges.txt").write_text("#version: 0.2\n")
    return CLIPTokenizer(str(d / "vocab.json"), str(d / "merges.txt"), model_max_length=77), len(vocab)


@pytest.fixture()
def patched(monkeypatch):
    """The native CLIP towers on the CPU stand-in ops (fp32); the kernels are exercised by tests/test_clip_gpu.py."""
    import fake_ops
    import imagharmony_b200.clip as clip
    import imagharmony_b200.ops as real_ops
    for name in dir(fake_ops):
        if not name.startswith("_") and callable(getattr(fake_ops, name)) and hasattr(real_ops, name):
            monkeypatch.setattr(real_ops, name, getattr(fake_ops, name))
    monkeypatch.setattr(clip, "_DTYPE", [torch.float32])
    yield


def test_clip_prompt_encoder_semantics(tmp_path, patched):
    transformers = pytest.importorskip("transformers")
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    from ip_adapter.encoders import ClipPromptEncoder
    tok1, nv = _toy_tokenizer(tmp_path, "t1")
    tok2, _ = _toy_tokenizer(tmp_path, "t2")
    torch.manual_seed(0)
    common = dict(vocab_size=nv, max_position_embeddings=77, num_hidden_layers=3, num_attention_heads=2,
                  bos_token_id=nv - 2, eos_token_id=nv - 1, pad_token_id=nv - 1)
    e1 = CLIPTextModel(CLIPTextConfig(hidden_size=32, intermediate_size=64, **common))
    e2 = CLIPTextModelWithProjection(CLIPTextConfig(hidden_size=48, intermediate_size=96, projection_dim=40, **common))
    enc = ClipPromptEncoder(tok1, tok2, e1, e2, device="cpu", dtype=torch.float32)
    prompts = ["eight sheep", "lions, best quality"]
    hs, pooled = enc(prompts)
    assert hs.shape == (2, 77, 32 + 48) and pooled.shape == (2, 40)
    assert hs.dtype == torch.float16 and pooled.dtype == torch.float16
    # spelled out: penultimate hidden states of both encoders, pooled = projected text_embeds of the second one
    with torch.no_grad():
        ids1 = tok1(prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        ids2 = tok2(prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        o1 = e1(ids1, output_hidden_states=True)
        o2 = e2(ids2, output_hidden_states=True)
    want = torch.cat([o1.hidden_states[-2], o2.hidden_states[-2]], dim=-1)
    assert torch.allclose(hs.float(), want, atol=2e-3)
    assert torch.allclose(pooled.float(), o2.text_embeds, atol=2e-3)
    assert ids1.shape == (2, 77) and int(ids1[0, 0]) == nv - 2                 # <|startoftext|> ... padded to 77


def test_pipeline_zero_negative_embeds_when_no_negative_prompt():
    """[3P] force_zeros_for_empty_prompt (SDXL-base): negative_prompt=None -> zero embeddings, a string -> encoded."""
    from imagharmony_b200.config import TINY
    from ip_adapter.custom_pipelines import StableDiffusionXLCustomPipeline
    pipe = StableDiffusionXLCustomPipeline.from_random(TINY, seed=0, device="cpu")
    pe, ne, pp, npool = pipe.encode_prompt("lions", num_images_per_prompt=2, do_classifier_free_guidance=True)
    assert pe.shape == (2, 77, TINY.cross_attention_dim) and torch.count_nonzero(ne) == 0 and torch.count_nonzero(npool) == 0
    pe2, ne2, _, npool2 = pipe.encode_prompt("lions", num_images_per_prompt=2, do_classifier_free_guidance=True,
                                             negative_prompt="blurry")
    assert torch.equal(pe, pe2) and torch.count_nonzero(ne2) > 0 and torch.count_nonzero(npool2) > 0
