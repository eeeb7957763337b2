"""Worker of tests/test_pns_cpu.py: one rank of a 2-process gloo group; writes its PNS result to a file."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagharmony_b200.pns import LinearProbeScorer, pns_select, pns_two_phase  # noqa: E402


def fake_runner(seeds):
    return torch.cat([torch.randn((1, 4, 8, 8), generator=torch.Generator("cpu").manual_seed(int(s))) for s in seeds]).half()


def fake_rest(preview):
    """stand-in for "denoise the winner to the end": a deterministic function of the preview latent"""
    return (preview.float() * 0.5 - 0.25).half()


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    seeds = [int(s) for s in sys.argv[5].split(",")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if len(sys.argv) > 6 and sys.argv[6] == "two_phase":
        res = pns_two_phase(fake_runner, fake_rest, seeds, LinearProbeScorer(4 * 8 * 8, seed=5), dist=dist, max_batch=2)
    else:
        res = pns_select(fake_runner, seeds, LinearProbeScorer(4 * 8 * 8, seed=5), dist=dist, max_batch=2)
    torch.save({"scores": res.scores, "best_index": res.best_index, "best_seed": res.best_seed,
                "best_latents": res.best_latents}, out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
