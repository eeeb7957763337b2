"""PNS shard / score all_gather / argmax on two CPU processes over gloo (the N>1 host logic of bench.py --pns)."""
import os
import socket
import subprocess
import sys

import torch

from imagharmony_b200.pns import LinearProbeScorer, pns_select, pns_two_phase, shard_seeds

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from pns_worker import fake_rest, fake_runner  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_seeds_covers_everything_once():
    seeds = list(range(10))
    for world in (1, 2, 3, 4, 8, 16):
        parts = [shard_seeds(seeds, r, world) for r in range(world)]
        assert sum(parts, []) == seeds
    assert shard_seeds([1, 2, 3], 5, 8) == []          # a rank with zero candidates


def test_pns_two_ranks_gloo_matches_single_process(tmp_path):
    seeds = [11, 12, 13, 14, 15]                        # uneven shards: 3 + 2
    single = pns_select(fake_runner, seeds, LinearProbeScorer(4 * 8 * 8, seed=5), dist=None, max_batch=2)
    port = str(_free_port())
    outs = [str(tmp_path / f"r{r}.pt") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "pns_worker.py"), str(r), "2", port, outs[r],
                               ",".join(map(str, seeds))]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    for o in outs:
        got = torch.load(o)
        assert torch.allclose(got["scores"], single.scores, atol=1e-5)
        assert got["best_index"] == single.best_index and got["best_seed"] == single.best_seed
        assert torch.equal(got["best_latents"], single.best_latents)   # every rank ends up with the winner's latents


def test_pns_two_phase_two_ranks_gloo(tmp_path):
    """preview all candidates -> all_gather scores -> broadcast the winning preview -> every rank finishes the winner"""
    seeds = [21, 22, 23, 24, 25, 26, 27]                # 4 + 3
    single = pns_two_phase(fake_runner, fake_rest, seeds, LinearProbeScorer(4 * 8 * 8, seed=5), dist=None, max_batch=2)
    assert torch.equal(single.best_latents, fake_rest(fake_runner([single.best_seed]))[0])
    port = str(_free_port())
    outs = [str(tmp_path / f"t{r}.pt") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "pns_worker.py"), str(r), "2", port, outs[r],
                               ",".join(map(str, seeds)), "two_phase"]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    for o in outs:
        got = torch.load(o)
        assert torch.allclose(got["scores"], single.scores, atol=1e-5)
        assert got["best_seed"] == single.best_seed
        assert torch.equal(got["best_latents"], single.best_latents)
