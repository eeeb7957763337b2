"""Host-side helpers with the reference's names (ip_adapter/utils.py:80-93).  The attention-map hook utilities of the
reference file (:6-79) have no caller anywhere in the reference and are out of scope."""
import torch
import torch.nn.functional as F


def is_torch2_available():
    return hasattr(F, "scaled_dot_product_attention")


def get_generator(seed, device):
    """seed -> torch.Generator; a list of seeds -> a list of generators, one per image (utils.py:83-93).  One generator
    per image is what makes a PNS candidate noise identical on whichever rank / batch slot it lands."""
    def make(value):
        return torch.Generator(device).manual_seed(value)
    if seed is None:
        return None
    return [make(v) for v in seed] if isinstance(seed, list) else make(seed)
