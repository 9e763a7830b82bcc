"""CPU, world_size 2 over gloo: the N > 1 logic of bench.py -- per-rank episode shards (no data-path collective) and the
max-over-ranks timing reduction that turns per-rank times into the whole-job nav-steps/s."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from gridmm_b200 import synth
    # each rank builds its own shard; shards differ and are reproducible
    ep = synth.make_episodes(2, 2, seed=bench.shard_seed(rank), dim=8)
    digest = torch.tensor([float(ep["depth_sub"].astype(np.int64).sum()), float(ep["pos"].sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    # rank r pretends its timed region took (10 + 5 r) ms for 7 steps
    ms, value = bench.aggregate(10.0 + 5.0 * rank, 7, world)
    out[rank] = (ms, value, [g.tolist() for g in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    import bench
    for r in range(world):
        ms, value, digests = res[r]
        assert ms == 15.0                                            # MAX over ranks, identical on every rank
        assert abs(value - world * bench.B * 7 / 15e-3) < 1e-6       # whole-job aggregate: all ranks' units / max time
        assert digests[0] != digests[1]                              # different shards per rank
    assert res[0][2] == res[1][2]
