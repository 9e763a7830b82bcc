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


def _grad_worker(rank, world, port, out):
    """Flat parameter / gradient buffers + bucketed all-reduce driven by autograd hooks (gridmm_b200.train), on CPU tensors."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gridmm_b200.train import FlatParams, GradientStep
    torch.manual_seed(0)                                   # same initial replica on every rank
    net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.LayerNorm(40), torch.nn.GELU(), torch.nn.Linear(40, 24),
                              torch.nn.LayerNorm(24), torch.nn.Linear(24, 3))
    ref = [p.detach().clone() for p in net.parameters()]
    flat = FlatParams(net)
    # values survive the re-homing, parameters are views of the flat buffer, the no-decay group sits behind the decayed one
    same = all(torch.equal(a, b) for a, b in zip(ref, net.parameters()))
    views = all(p.data_ptr() >= flat.params.data_ptr() and p.grad.data_ptr() >= flat.grads.data_ptr() for p in net.parameters())
    names = [n for n, _ in flat.order]
    n_dec = sum(1 for n in names if not (n.endswith("bias") or "LayerNorm" in n))
    group_ok = all(not n.endswith("bias") for n in names[:n_dec]) and all(n.endswith("bias") for n in names[n_dec:])
    gs = GradientStep(flat, lr=1e-3, bucket_elems=1200)    # several buckets
    torch.manual_seed(100 + rank)                          # different data per rank
    x, y = torch.randn(16, 24), torch.randn(16, 3)
    gs.arm()
    loss = ((net(x) - y) ** 2).mean()
    loss.backward()
    gs.reduce_all()
    got = flat.grads.clone()
    # reference: the same backward on a plain copy, gradients summed over the ranks tensor by tensor
    torch.manual_seed(0)
    net2 = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.LayerNorm(40), torch.nn.GELU(), torch.nn.Linear(40, 24),
                               torch.nn.LayerNorm(24), torch.nn.Linear(24, 3))
    ((net2(x) - y) ** 2).mean().backward()
    want = {}
    for n, p in net2.named_parameters():
        g = p.grad.clone()
        dist.all_reduce(g)
        want[n] = g
    err = 0.0
    for (n, p), off in zip(flat.order, flat.offsets):
        err = max(err, (got[off:off + p.numel()].view(p.shape) - want[n]).abs().max().item())
    out[rank] = (same, views, group_ok, len(gs.buckets), err, got.sum().item())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_buckets():
    """The communication half of the pretraining gradient step: gradients land in one flat buffer, buckets are
    all-reduced from autograd hooks while backward is still running, and the result equals a per-tensor all-reduce."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for r in range(world):
        same, views, group_ok, n_buckets, err, _ = res[r]
        assert same and views and group_ok
        assert n_buckets >= 3
        assert err < 1e-6
    assert abs(res[0][5] - res[1][5]) < 1e-5                # both ranks hold the same reduced gradients
