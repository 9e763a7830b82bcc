"""The gradient step of distributed pretraining (BASELINE config 5): what happens between `loss.backward()` and the
next forward in the reference's loop (pretrain_src/train_r2r.py:258-296, DDP wrapper pretrain_src/utils/misc.py:57-58):

    gradient all-reduce over the ranks  ->  clip_grad_norm_  ->  AdamW (two parameter groups)  ->  zero_grad

B200-native form:
  * every parameter and its gradient are VIEWS into two flat fp32 buffers (decayed group first, the bias / LayerNorm group
    behind it: pretrain_src/optim/misc.py:12-22), so the all-reduce is a handful of large NCCL calls over contiguous
    memory (NVLink 5 / NVSwitch: bucket size is chosen for launch latency and overlap, not link count) instead of one per tensor;
  * buckets are reduced as soon as the autograd engine has produced their gradients (post-accumulate hooks; parameters are laid
    out in REVERSE registration order so that the gradients that arrive first fill the first bucket), overlapping NCCL with the
    rest of backward;
  * the global gradient norm, the clip factor and the AdamW update run in three launches of this package's kernels over the flat
    buffers (gridmm_grad_sumsq / gridmm_adamw_step): the clip factor never visits the host.

This module is the optimizer / communication half only.  The kernels behind forward('navigation') and forward_pretrain are
inference kernels (no backward); `GradientStep` works with any autograd graph whose leaves are the flattened parameters.
"""
import torch
import torch.distributed as dist

from . import ops

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")        # pretrain_src/optim/misc.py:13


def _align(n, a=64):
    return (n + a - 1) // a * a


class FlatParams:
    """Re-homes the parameters of `module` into one flat fp32 buffer (and their .grad into another), group by group."""

    def __init__(self, module, no_decay=NO_DECAY):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("no trainable parameters")
        dev = named[0][1].device
        decay = [(n, p) for n, p in named if not any(nd in n for nd in no_decay)]
        nodec = [(n, p) for n, p in named if any(nd in n for nd in no_decay)]
        # reverse registration order inside each group: autograd reaches the last layers first
        self.order = list(reversed(decay)) + list(reversed(nodec))
        sizes = [_align(p.numel()) for _, p in self.order]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.n_decay = sum(sizes[:len(decay)])
        total = self.offsets[-1]
        self.params = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for (name, p), off in zip(self.order, self.offsets):
                n = p.numel()
                self.params[off:off + n].copy_(p.detach().reshape(-1).float())
                p.data = self.params[off:off + n].view(p.shape)
                p.grad = self.grads[off:off + n].view(p.shape)
        self.total = total

    def zero_grad(self):
        self.grads.zero_()


class GradientStep:
    """All-reduce (bucketed, overlapped with backward) + clip + AdamW over a FlatParams.

        flat = FlatParams(model); gs = GradientStep(flat, lr=5e-5, weight_decay=0.01, max_norm=5.0)
        gs.arm(); loss.backward(); gs.step()          # every iteration
    """

    def __init__(self, flat, lr, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_norm=-1.0, bucket_elems=32 << 20, group=None, after_step=()):
        self.flat, self.lr, self.betas, self.eps, self.wd, self.max_norm = flat, lr, betas, eps, weight_decay, max_norm
        self.group = group
        self.after_step = list(after_step)      # callables run after every update (e.g. a weight cache's invalidate)
        self._last_scale = 1.0
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.m = torch.zeros_like(flat.params)
        self.v = torch.zeros_like(flat.params)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=flat.params.device)
        self.t = 0
        # buckets: contiguous ranges of the flat gradient buffer, cut at parameter boundaries
        self.buckets = []            # (start, end, [indices of the parameters inside])
        start, members = 0, []
        for i, ((name, p), off) in enumerate(zip(flat.order, flat.offsets)):
            members.append(i)
            end = flat.offsets[i + 1]
            if end - start >= bucket_elems or i == len(flat.order) - 1:
                self.buckets.append((start, end, members))
                start, members = end, []
        self._bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self._bucket_of[i] = b
        self._pending = None
        self._works = []
        self._hooks = []
        if self.world > 1:
            for i, (name, p) in enumerate(flat.order):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    def _make_hook(self, i):
        def hook(param):
            if self._pending is None:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._reduce_bucket(b)
        return hook

    def _reduce_bucket(self, b):
        s, e, _ = self.buckets[b]
        # SUM now, the 1 / world average is folded into the update kernel's grad_scale
        self._works.append(dist.all_reduce(self.flat.grads[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def arm(self):
        """Call before backward: buckets are all-reduced as soon as every gradient inside them has been accumulated."""
        self._pending = [len(mem) for _, _, mem in self.buckets] if self.world > 1 else None
        self._works = []

    def reduce_all(self):
        """All-reduce every bucket that the hooks have not launched (no autograd graph, or parameters without a gradient)."""
        if self.world > 1:
            if self._pending is None:
                self._pending = [1] * len(self.buckets)
            for b, left in enumerate(self._pending):
                if left > 0:
                    self._reduce_bucket(b)
            for w in self._works:
                w.wait()
        self._pending = None
        self._works = []

    def step(self, lr=None, loss_scale=1.0):
        """clip_grad_norm_ + AdamW over both groups (train_r2r.py:281-296), then zero_grad."""
        self.reduce_all()
        self.t += 1
        lr = self.lr if lr is None else lr
        f = self.flat
        scale = 1.0 / (self.world * loss_scale)
        sumsq = None
        if self.max_norm is not None and self.max_norm > 0:
            self.sumsq.zero_()
            ops.grad_sumsq(f.grads, self.sumsq)
            sumsq = self.sumsq
        nd = f.n_decay
        if nd > 0:
            ops.adamw_step(f.params[:nd], f.grads[:nd], self.m[:nd], self.v[:nd], lr, self.betas[0], self.betas[1], self.eps, self.wd,
                           self.t, grad_scale=scale, sumsq=sumsq, max_norm=self.max_norm or 0.0)
        if f.total > nd:
            ops.adamw_step(f.params[nd:], f.grads[nd:], self.m[nd:], self.v[nd:], lr, self.betas[0], self.betas[1], self.eps, 0.0,
                           self.t, grad_scale=scale, sumsq=sumsq, max_norm=self.max_norm or 0.0)
        f.zero_grad()
        self._last_scale = scale
        for cb in self.after_step:
            cb()

    def grad_norm(self):
        """Global gradient norm of the last step (after averaging), as clip_grad_norm_ returns it.  Host synchronisation."""
        return float(self.sumsq.sqrt().item()) * self._last_scale
