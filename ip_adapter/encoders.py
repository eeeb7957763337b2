"""SDXL prompt conditioning through the two CLIP text encoders (scope row f2, [3P] diffusers `encode_prompt` as called at
ip_adapter.py:292-297,314-319 / custom_pipelines.py:229-247).

The encoders run ONCE per generate() and are outside the denoise hot path and the benchmark metric, so
they use the `transformers` library classes the reference itself uses (CLIPTextModel, CLIPTextModelWithProjection,
CLIPTokenizer) rather than native kernels.  Semantics restated from diffusers==0.30.0:
  * each prompt is tokenised by both tokenizers with padding="max_length" (77) and truncation;
  * prompt_embeds = concat(hidden_states[-2] of encoder 1 [.., 768], hidden_states[-2] of encoder 2 [.., 1280]) -> 2048;
  * pooled_prompt_embeds = `text_embeds` of encoder 2 (the projected pooled output) -> 1280.
No weights or tokenizer vocabularies exist offline; tests build miniature random CLIP models and a toy vocabulary."""
from __future__ import annotations

import os
from typing import List, Tuple

import torch


class ClipPromptEncoder:
    def __init__(self, tokenizer, tokenizer_2, text_encoder, text_encoder_2, device="cuda", dtype=torch.float16):
        self.tokenizers = [tokenizer, tokenizer_2]
        self.text_encoders = [text_encoder.to(device=device, dtype=dtype).eval(),
                              text_encoder_2.to(device=device, dtype=dtype).eval()]
        self.device, self.dtype = device, dtype

    @classmethod
    def from_pretrained(cls, path: str, device="cuda", dtype=torch.float16) -> "ClipPromptEncoder":
        """`path` = an SDXL snapshot folder with tokenizer/, tokenizer_2/, text_encoder/, text_encoder_2/."""
        from transformers import CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer
        return cls(CLIPTokenizer.from_pretrained(os.path.join(path, "tokenizer")),
                   CLIPTokenizer.from_pretrained(os.path.join(path, "tokenizer_2")),
                   CLIPTextModel.from_pretrained(os.path.join(path, "text_encoder")),
                   CLIPTextModelWithProjection.from_pretrained(os.path.join(path, "text_encoder_2")), device, dtype)

    @staticmethod
    def available(path: str) -> bool:
        return all(os.path.isdir(os.path.join(path, d)) for d in ("tokenizer", "tokenizer_2", "text_encoder",
                                                                    "text_encoder_2"))

    @torch.no_grad()
    def __call__(self, prompts: List[str]) -> Tuple[torch.Tensor, torch.Tensor]:
        embeds, pooled = [], None
        for tok, enc in zip(self.tokenizers, self.text_encoders):
            ids = tok(list(prompts), padding="max_length", max_length=tok.model_max_length, truncation=True,
                      return_tensors="pt").input_ids.to(self.device)
            out = enc(ids, output_hidden_states=True)
            pooled = out[0]                            # kept from the LAST encoder only: its projected `text_embeds`
            embeds.append(out.hidden_states[-2])       # penultimate layer (clip_skip = None)
        return torch.cat(embeds, dim=-1).to(torch.float16), pooled.to(torch.float16)
