"""PNS shard / score all_gather / argmax on two CPU processes over gloo (the N>1 host logic of bench.py --pns)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from imagharmony_b200.pns import LinearProbeScorer, pns_select, shard_seeds


def _fake_runner(seeds):
    # stands in for DenoiseEngine.run: latents are a deterministic function of the seed only (placement invariant)
    return torch.cat([torch.randn((1, 4, 8, 8), generator=torch.Generator("cpu").manual_seed(int(s))) for s in seeds]).half()


def _worker(rank, world, port, seeds, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = pns_select(_fake_runner, seeds, LinearProbeScorer(4 * 8 * 8, seed=5), dist=dist, max_batch=2)
    q.put((rank, res.scores.clone(), res.best_index, res.best_seed, res.best_latents.clone()))
    dist.barrier()
    dist.destroy_process_group()


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


def test_pns_two_ranks_gloo_matches_single_process():
    seeds = [11, 12, 13, 14, 15]                        # uneven shards: 3 + 2
    single = pns_select(_fake_runner, seeds, LinearProbeScorer(4 * 8 * 8, seed=5), dist=None, max_batch=2)
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seeds, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scores, bi, bs, lat in got:
        assert torch.allclose(scores, single.scores, atol=1e-5)
        assert bi == single.best_index and bs == single.best_seed
        assert torch.equal(lat, single.best_latents)      # every rank ends up with the winner's latents
