"""BASELINE config 5: one optimizer step of the R2R pretraining loop (MLM + SAP proxy tasks, batch 32 per GPU, DDP-style gradient
averaging over the GPUs of one box), i.e. what pretrain_src/train_r2r.py:229-296 does per iteration:

    batch -> GPU, forward(task), loss.backward()  [gradient all-reduce overlapped], clip_grad_norm_, AdamW, zero_grad

with gridmm_b200.train_model.PretrainModel (every nn.Linear forward / dgrad / wgrad on this package's tcgen05 GEMM; LayerNorm,
attention, pooling and the losses are torch ops) and gridmm_b200.train.GradientStep (flat fp32 buckets all-reduced from autograd
hooks over NCCL, global-norm clip and AdamW in this package's kernels).

    python tools/bench_pretrain.py --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/bench_pretrain.py --gpus 8 --steps 20 --warmup 3
    python tools/bench_pretrain.py --impl reference --steps 1 --warmup 0        # the same step by CPU autograd through oracle/

One JSON line (bench.py's contract; `bench.py --workload pretrain` forwards here).  The tasks alternate mlm / sap step by step (the
reference samples them 1:1, r2r_pretrain.json mix_ratio), synthetic batches of the dataset's shape: paths of 2..8 viewpoints x 36
views, 80-token instructions, 588 grid points per viewpoint.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gridmm_b200 import synth  # noqa: E402

B_PER_GPU, TXT_LEN, MAX_STEPS, N_MASKED = 32, 80, 8, 8
CONFIG = {"workload": "configs[4]: R2R pretrain (MLM+SAP proxy tasks), batch %d/GPU (global 256 at 8 GPUs), paths of 2..%d viewpoints x 36 "
                      "views, %d-token instructions, 588 grid points / viewpoint; one step = forward + backward + gradient all-reduce + "
                      "clip + AdamW, tasks alternating mlm / sap" % (B_PER_GPU, MAX_STEPS, TXT_LEN),
          "batch_per_gpu": B_PER_GPU, "parallelism": "data parallel: one process per GPU, flat-bucket gradient all-reduce (NCCL) overlapped "
                                                     "with backward", "dropout": 0.1, "loss_scale": 1024.0}


def model_kwargs():
    return dict(num_l_layers=9, num_pano_layers=2, num_x_layers=4)          # pretrain_src/config/r2r_model_config.json


def make_batch(seed, batch=B_PER_GPU):
    """A collated pretraining batch on the host (pinned where it is a tensor), grid tensors of the dataset's shape."""
    pb = synth.make_pretrain_batch(batch, seed=seed, txt_len=TXT_LEN, max_steps=MAX_STEPS)
    out = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    for k, v in synth.make_pretrain_labels(pb, seed=seed, n_masked=N_MASKED).items():
        out[k] = torch.from_numpy(v)
    rng = np.random.default_rng(seed + 7)
    gf, gm = [], []
    for b in range(batch):
        n = 588 * int(pb["traj_step_lens"][b])
        gf.append(torch.from_numpy(rng.standard_normal((n, 768), dtype=np.float32).astype(np.float16)))
        cells = rng.integers(0, 196, size=n).astype(np.int64)
        cells[rng.random(n) < 0.05] = -1                                        # points outside the 14 x 14 window
        gm.append(torch.from_numpy(cells))
    out.update(grid_fts=gf, grid_map=gm, gridmap_pos_fts=torch.from_numpy(rng.standard_normal((batch, 196, 5), dtype=np.float32)))
    return out


def to_device(batch, dev, pin=False):
    def mv(v):
        if torch.is_tensor(v):
            return v.pin_memory() if pin else v.to(dev)
        if isinstance(v, list) and v and torch.is_tensor(v[0]):
            return [mv(x) for x in v]
        return v
    return {k: mv(v) for k, v in batch.items()}


def batch_bytes(batch):
    n = 0
    for v in batch.values():
        if torch.is_tensor(v):
            n += v.numel() * v.element_size()
        elif isinstance(v, list) and v and torch.is_tensor(v[0]):
            n += sum(x.numel() * x.element_size() for x in v)
    return n


def build_model(seed=0):
    from gridmm_b200.model import NavConfig
    from gridmm_b200.train_model import PretrainModel
    torch.manual_seed(seed)                            # random-init weights of the reference's architecture (normal(0, 0.02))
    return PretrainModel(NavConfig(pretrain_trunk=True, use_lang2visn_attn=True, graph_sprels=False, **model_kwargs()))


def run_ours(args):
    import torch.distributed as dist
    from gridmm_b200.train import FlatParams, GradientStep
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench_pretrain: no CUDA device (the training step has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(0).to(dev).train()
    flat = FlatParams(model)
    scale = CONFIG["loss_scale"]
    gs = GradientStep(flat, lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_norm=5.0, after_step=[model.weights_updated])
    host = [make_batch(1000 * rank + i) for i in range(4)]
    pinned = [to_device(b, dev, pin=True) for b in host]
    resident = [to_device(b, dev) for b in host]
    tasks = ["mlm", "sap"]
    state = {"i": 0, "loss": None}

    def step(batches):
        i = state["i"]
        state["i"] += 1
        gs.arm()
        loss = model(batches[i % len(batches)], tasks[i % 2]).mean()
        (loss * scale).backward()
        gs.step(loss_scale=scale)
        state["loss"] = loss.detach()

    def timed(batches, steps, warmup, read_loss):
        for _ in range(warmup):
            step(batches)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(batches)
            if read_loss:
                state["loss"].item()                      # the loop logs the loss every step (train_r2r.py:262-266): D2H read
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    from bench import ClockSampler                        # noqa: E402  (same clocks line as the headline bench)
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.time()
    ms_res = timed(resident, args.steps, args.warmup, False)
    clocks = sampler.stop(t0, time.time()) if sampler is not None else None
    ms_e2e = timed(pinned, args.steps, max(1, args.warmup // 2), True)

    # where the time goes (rank 0, synchronised phases; explains the number, is not the number)
    phases = {}
    if rank == 0 or world > 1:
        def phase():
            t = {}
            torch.cuda.synchronize(); a = time.perf_counter()
            gs.arm()
            loss = model(resident[0], tasks[state["i"] % 2]).mean()
            torch.cuda.synchronize(); b = time.perf_counter()
            (loss * scale).backward()
            torch.cuda.synchronize(); c = time.perf_counter()
            gs.step(loss_scale=scale)
            torch.cuda.synchronize(); d = time.perf_counter()
            state["i"] += 1
            return (b - a) * 1e3, (c - b) * 1e3, (d - c) * 1e3
        rows = np.array([phase() for _ in range(4)])
        phases = {"forward_ms": float(rows[:, 0].mean()), "backward_ms(incl. overlapped all-reduce launches)": float(rows[:, 1].mean()),
                  "reduce_wait+clip+adamw_ms": float(rows[:, 2].mean())}
    if rank == 0:
        n_params = sum(p.numel() for p in model.parameters())
        line = {"metric": "pretrain samples/sec (MLM+SAP proxy step)", "value": world * B_PER_GPU / (ms_res / 1e3), "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp16 GEMM operands, fp32 accumulate / master weights / optimizer",
                "data": "synthetic", "config": dict(CONFIG, parameters=n_params, gradient_bytes_all_reduced=4 * flat.total if world > 1 else 0),
                "e2e": {"value": world * B_PER_GPU / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": batch_bytes(host[0]), "d2h_bytes_per_step": 4,
                        "mode": "batch collated in pinned host memory, copied inside the step; the loss is read back every step"},
                "phases": phases, "clocks": clocks,
                "note": "LinearFn GEMMs (forward, dgrad, wgrad), cast/transpose, column sums, gradient norm and AdamW are this package's "
                        "kernels; LayerNorm, attention, pooling, losses and their backward are torch ops in this round"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The same step by CPU autograd through the oracle restatement of the reference's wrapper + torch AdamW, all host cores."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle import pretrain_oracle as po
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_model(0)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items() if k != "mlm_head.predictions.decoder.weight"}
    sd["mlm_head.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    params = [v for k, v in sd.items() if k != "mlm_head.predictions.decoder.weight"]
    opt = torch.optim.AdamW(params, lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    kw = dict(n_l_layers=9, n_pano_layers=2, n_x_layers=4)
    batches = [make_batch(i) for i in range(2)]

    def step(i):
        b = batches[i % 2]
        if i % 2 == 0:
            scores = po.mlm_scores(sd, b, b["txt_labels"], **kw)
            loss = torch.nn.functional.cross_entropy(scores, b["txt_labels"][b["txt_labels"] != -1])
        else:
            labels = {k: b[k] for k in ("gmap_visited_masks", "global_act_labels", "local_act_labels")}
            loss = po.sap(sd, b, labels, **kw)[3].mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        opt.zero_grad()
    for i in range(args.warmup):
        step(i)
    steps = max(2, min(args.steps, 4))
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    v = B_PER_GPU / (ms / 1e3)
    print(json.dumps({"impl": "reference", "metric": "pretrain samples/sec (MLM+SAP proxy step)", "value": v, "unit": "samples/s",
                      "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
                      "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                       "sample": "%d optimizer steps (mlm, sap alternating) of one rank's batch of 32" % steps},
                      "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args, _ = ap.parse_known_args(argv)
    (run_reference if args.impl == "reference" else run_ours)(args)


if __name__ == "__main__":
    main()
