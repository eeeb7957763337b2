"""The reference ships a byte-identical copy of `shared_models.py` inside the package; same re-export here."""
from imagharmony_b200.adapter import ImageProjModel  # noqa: F401
