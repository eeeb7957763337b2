"""The reference imports `HarmonyAttention` (ip_adapter/ip_adapter.py:10) and `ComposedAttention` (demo.py:11) from a
module of this name that it does not ship; this shim makes those imports resolve to the native modules."""
from imagharmony_b200.adapter import ComposedAttention, HarmonyAttention  # noqa: F401
