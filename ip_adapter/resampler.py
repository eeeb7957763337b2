"""Resampler with the reference's import path (ip_adapter/resampler.py:81-147); implementation on the sm_100a kernels
lives in imagharmony_b200/adapter.py."""
from imagharmony_b200.adapter import PerceiverAttention, Resampler, _FeedForward  # noqa: F401


def FeedForward(dim, mult=4):
    """resampler.py:13-20: LayerNorm, Linear(dim -> dim*mult, no bias), GELU, Linear(back, no bias) with the reference's
    Sequential indices (state-dict keys `0.weight, 0.bias, 1.weight, 3.weight`), running on the native kernels."""
    return _FeedForward(dim, mult)


def reshape_tensor(x, heads):
    """[bs, length, heads * d] -> [bs, heads, length, d] (resampler.py:23-31); host-side layout helper -- the native
    PerceiverAttention never materialises this permutation (the attention kernel reads heads as column blocks)."""
    bs, length, _ = x.shape
    return x.reshape(bs, length, heads, -1).permute(0, 2, 1, 3).reshape(bs, heads, length, -1)


def masked_mean(t, *, dim, mask=None):
    """resampler.py:150-158 (host-side utility; the native Resampler uses ops.mean_tokens for the all-ones mask)."""
    if mask is None:
        return t.mean(dim=dim)
    denom = mask.sum(dim=dim, keepdim=True)
    masked_t = t.masked_fill(~mask.unsqueeze(-1), 0.0)
    return masked_t.sum(dim=dim) / denom.clamp(min=1e-5)
