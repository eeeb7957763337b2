"""`from shared_models import ImageProjModel` (train.py:23 of the reference) resolves to the native module.  The rest of
the reference file of this name -- an older, caller-less `Composed_Attention` and its own `Cross_Attention` variant
(shared_models.py:16-63, 88-171) -- is out of scope."""
from imagharmony_b200.adapter import ImageProjModel  # noqa: F401
