"""test.py:5 of the reference does `from train import HarmonyAttention`; training itself is out of scope for this
inference hot path, so this module only re-exports the inference-time modules."""
from imagharmony_b200.adapter import HarmonyAttention, ImageProjModel  # noqa: F401
