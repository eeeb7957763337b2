"""Training checkpoint -> `ip_adapter.bin` (scope row f3).  Same module name, function name, arguments and return
convention as the reference tool (convert_bin.py:5-49): the flat state dict that accelerate saves for the composite
training module (prefixes `image_proj_model.`, `adapter_modules.`, `composed_modules.`; train.py's IPAdapter wrapper) is
regrouped into the 3-key dict `{"image_proj", "ip_adapter", "composed_adapter"}` that `IPAdapterXL.load_ip_adapter`
reads.  Everything else in the checkpoint (the frozen UNet) is dropped.

    python convert_bin.py <pytorch_model.bin> [<ip_adapter.bin>]      # default output: next to the input

The reference's `__main__` walks a hard-coded directory of `checkpoint-*` folders; here the paths come from the command
line, and a directory argument converts every `checkpoint-*/pytorch_model.bin` below it.
"""
import glob
import os
import sys

import torch

from imagharmony_b200.weights import split_ip_adapter_checkpoint


def convert_checkpoint_to_ip_adapter(pytorch_model_path, output_ip_adapter_path):
    """True when `output_ip_adapter_path` was written.  A missing source, a checkpoint without any of the three
    prefixes, or an unreadable file give a message and False -- no exception, no output file (convert_bin.py:7-9, 31-33,
    45-47)."""
    if not os.path.exists(pytorch_model_path):
        print(f"  [Warning] Source file not found, skipping: {pytorch_model_path}")
        return False
    print(f"  Converting: {pytorch_model_path}")
    try:
        flat = torch.load(pytorch_model_path, map_location="cpu")
        parts = split_ip_adapter_checkpoint(flat)
        if not any(parts.values()):
            print("  [Warning] No expected keys (image_proj_model, adapter_modules, composed_modules) found in "
                  f"{pytorch_model_path}. Skipping save.")
            return False
        torch.save(parts, output_ip_adapter_path)
    except Exception as ex:  # the reference reports and carries on with the next checkpoint
        print(f"  [Error] Failed to convert {pytorch_model_path}: {ex}")
        return False
    print(f"  Successfully saved: {output_ip_adapter_path}")
    return True


def _main(argv):
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    src = argv[0]
    if os.path.isdir(src):
        jobs = [(p, os.path.join(os.path.dirname(p), "ip_adapter.bin"))
                for p in sorted(glob.glob(os.path.join(src, "checkpoint-*", "pytorch_model.bin")))]
    else:
        jobs = [(src, argv[1] if len(argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(src)), "ip_adapter.bin"))]
    done = sum(bool(convert_checkpoint_to_ip_adapter(a, b)) for a, b in jobs)
    print(f"converted {done} of {len(jobs)} checkpoint(s)")
    return 0 if done == len(jobs) and jobs else 1


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
