"""Gradient step of distributed pretraining (BASELINE config 5: global batch 256 over 8 x B200, DDP): all-reduce of the flat
gradient buffer over NCCL / NVLink + global-norm clip + AdamW, for the parameter tree of the reference's pretraining model
(pretrain_src/model/vilmodel.py:640-666 trunk, 9 language / 2 panorama / 4 cross-modal layers, lang2visn blocks).

    python tools/bench_gradstep.py                                   # 1 GPU: the update kernels against the HBM roofline
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_gradstep.py

This is the optimizer / communication half only (gridmm_b200/train.py); backward is not part of this package.  Prints one JSON
line on rank 0: ms per gradient step (max over ranks, CUDA events), the all-reduce's algorithmic bus bandwidth and the update
kernels' achieved HBM bandwidth (28 B per parameter: p, g, m, v read; p, m, v written; + 4 B for the norm)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridmm_b200.model import GlocalTextPathNavCMT, NavConfig       # noqa: E402
from gridmm_b200.train import FlatParams, GradientStep              # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = NavConfig(pretrain_trunk=True, use_lang2visn_attn=True, num_l_layers=9, num_pano_layers=2, num_x_layers=4)
    model = GlocalTextPathNavCMT(cfg).to(dev)
    flat = FlatParams(model)
    gs = GradientStep(flat, lr=5e-5, weight_decay=0.01, max_norm=5.0)
    n = flat.total
    steps, warmup = 20, 5
    g = torch.Generator(device=dev).manual_seed(rank)

    def one(comm=True):
        flat.grads.normal_(0.0, 1e-3, generator=g) if False else None      # gradients stay as they are (content does not matter)
        gs.arm()
        if comm:
            gs.reduce_all()
        else:
            gs._pending = None
        gs.step()

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    flat.grads.normal_(0.0, 1e-3, generator=g)
    ms_step = timed(lambda: one(True))
    # the update alone (no communication): same code with the all-reduce skipped
    saved_world = gs.world
    gs.world = 1
    ms_update = timed(lambda: one(False))
    gs.world = saved_world
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        upd_bytes = n * 4 * (4 + 3 + 1)            # p, g, m, v in; p, m, v out; g once more for the norm
        zero_bytes = n * 4
        ar_ms = max(ms_step - ms_update, 0.0)
        line = {"metric": "pretraining gradient step (flat all-reduce + clip + AdamW), ms", "n_gpus": world, "params": n,
                "ms_per_step": ms_step, "ms_update_only": ms_update, "ms_all_reduce_exposed": ar_ms,
                "update_gbs": (upd_bytes + zero_bytes) / (ms_update * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                "update_frac_of_hbm": (upd_bytes + zero_bytes) / (ms_update * 1e-3) / 1e9 / hbm,
                "allreduce_busbw_gbs": (2.0 * (world - 1) / world * n * 4 / (ar_ms * 1e-3) / 1e9) if world > 1 and ar_ms > 0 else None,
                "buckets": len(gs.buckets), "steps": steps, "warmup": warmup}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
