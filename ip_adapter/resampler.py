"""Resampler with the reference's import path (ip_adapter/resampler.py:81-147); implementation on the sm_100a kernels
lives in imagharmony_b200/adapter.py."""
from imagharmony_b200.adapter import PerceiverAttention, Resampler  # noqa: F401


def masked_mean(t, *, dim, mask=None):
    """resampler.py:150-158 (host-side utility; the native Resampler uses ops.mean_tokens for the all-ones mask)."""
    if mask is None:
        return t.mean(dim=dim)
    denom = mask.sum(dim=dim, keepdim=True)
    masked_t = t.masked_fill(~mask.unsqueeze(-1), 0.0)
    return masked_t.sum(dim=dim) / denom.clamp(min=1e-5)
