"""DiLoCo outer optimizer — distributed low-communication training (Douillard et al., arXiv:2311.08105), the training
regime the reference's README names as its goal and never implements.

Every *worker* — one data-parallel replica, which may itself be tensor- and pipeline-parallel — trains with its own
inner optimizer on its own data shard and exchanges **nothing** for ``inner_steps`` (H) steps.  Then the workers
average how far they moved, ``delta = theta_shared - theta_worker`` (the "outer gradient"), and an outer SGD with
Nesterov momentum applies it to the shared parameters, from which every worker continues:

    inner:  theta_k <- InnerOpt(theta_k, grad_k)                      H times, no communication
    outer:  g = mean_k(theta_shared - theta_k);  v = mu v + g;  theta_shared -= lr_out (g + mu v)   [Nesterov]

Communication drops from one gradient all-reduce per step to one parameter-sized all-reduce every H steps, which is
what makes poorly connected islands of well connected GPUs (NVLink inside a box, slow links between boxes) trainable.

Do NOT wrap the module in ``DataParallel`` when using this optimizer (that would re-introduce the per-step gradient
all-reduce); the DATA group of the :class:`ParallelContext` is the set of DiLoCo workers.  Tensor / pipeline
parallelism inside a worker work as usual — only parameters this rank owns are averaged, each over its DATA group.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode


class DiLoCoOptimizer:
    BUCKET_ELEMS = 16 * 1024 * 1024   # fp32 elements per outer all-reduce (64 MB): bounds the transient memory

    def __init__(self, inner_optim: torch.optim.Optimizer, parallel_context, inner_steps: int = 500, outer_lr: float = 0.7,
                 outer_momentum: float = 0.9, nesterov: bool = True, parallel_mode: ParallelMode = ParallelMode.DATA):
        assert inner_steps >= 1
        self.optim = inner_optim
        self.parallel_context = parallel_context
        self.parallel_mode = parallel_mode
        self.inner_steps = inner_steps
        self.outer_lr, self.outer_momentum, self.nesterov = outer_lr, outer_momentum, nesterov
        self.local_step = 0
        self.outer_step_count = 0
        self._anchor: Optional[List[torch.Tensor]] = None     # fp32 copy of the shared parameters
        self._momentum: Optional[List[torch.Tensor]] = None

    # ------------------------------------------------------------------ parameter access
    def _fused(self):
        from pipegoose_b200.optim.fused_adam import FusedAdam

        return self.optim if isinstance(self.optim, FusedAdam) else None

    def _tensors(self) -> List[torch.Tensor]:
        """The tensors that hold the authoritative parameter values: FusedAdam's fp32 master buffer (one flat
        tensor), otherwise every parameter's ``.data``."""
        fused = self._fused()
        if fused is not None:
            fused._lazy_init()
            assert fused._segments == [(0, fused.flat.numel)], "DiLoCo workers keep whole optimizer states (no ZeRO slices)"
            return [fused.master]
        return [p.data for g in self.optim.param_groups for p in g["params"] if p.numel() > 0]

    def _ensure_state(self):
        if self._anchor is None:
            self._anchor = [t.detach().float().clone() for t in self._tensors()]
            self._momentum = [torch.zeros_like(a) for a in self._anchor]

    # ------------------------------------------------------------------ optimizer API
    @property
    def param_groups(self):
        return self.optim.param_groups

    @property
    def defaults(self):
        return self.optim.defaults

    def add_param_group(self, *args, **kwargs):
        assert self._anchor is None, "parameter groups must be added before the first step"
        self.optim.add_param_group(*args, **kwargs)

    def zero_grad(self, *args, **kwargs):
        self._ensure_state()   # the anchor is the parameters BEFORE the first inner step
        self.optim.zero_grad(*args, **kwargs)

    @torch.no_grad()
    def step(self, *args, **kwargs):
        self._ensure_state()
        out = self.optim.step(*args, **kwargs)
        self.local_step += 1
        if self.local_step % self.inner_steps == 0:
            self.outer_step()
        return out

    @torch.no_grad()
    def outer_step(self):
        """Average the workers' displacement and move the shared parameters (collective over the DATA group)."""
        ctx, mode = self.parallel_context, self.parallel_mode
        n_workers = ctx.get_world_size(mode)
        tensors = self._tensors()
        deltas = [a - t.float() for a, t in zip(self._anchor, tensors)]
        if n_workers > 1 and deltas:
            group = ctx.get_group(mode)
            to_cuda = dist.get_backend(group) == "nccl"

            def average(tensors):   # one all-reduce per bucket; large tensors travel alone, in place
                flat = tensors[0].reshape(-1) if len(tensors) == 1 else torch.cat([t.reshape(-1) for t in tensors])
                buf = flat.cuda() if to_cuda and not flat.is_cuda else flat
                dist.all_reduce(buf, group=group)
                buf.div_(n_workers)
                if buf is not flat:
                    flat.copy_(buf)
                if len(tensors) > 1:
                    offset = 0
                    for t in tensors:
                        t.copy_(flat[offset:offset + t.numel()].view_as(t))
                        offset += t.numel()

            bucket, size = [], 0
            for d in deltas:
                if d.numel() >= self.BUCKET_ELEMS:
                    average([d])
                    continue
                bucket.append(d)
                size += d.numel()
                if size >= self.BUCKET_ELEMS:
                    average(bucket)
                    bucket, size = [], 0
            if bucket:
                average(bucket)
        mu = self.outer_momentum
        for anchor, vel, delta, tensor in zip(self._anchor, self._momentum, deltas, tensors):
            vel.mul_(mu).add_(delta)
            anchor.sub_(delta.add(vel, alpha=mu) if self.nesterov else vel, alpha=self.outer_lr)
            tensor.copy_(anchor.to(tensor.dtype))
        fused = self._fused()
        if fused is not None:   # the model reads the low-precision flat copy of the master weights
            fused.flat.flat_param.copy_(fused.master.to(fused.flat.dtype))
        self.outer_step_count += 1

    # ------------------------------------------------------------------ checkpoints
    def state_dict(self):
        self._ensure_state()
        return {"inner": self.optim.state_dict(), "anchor": self._anchor, "momentum": self._momentum,
                "local_step": self.local_step, "outer_step_count": self.outer_step_count}

    def load_state_dict(self, sd):
        self.optim.load_state_dict(sd["inner"])
        self._ensure_state()
        for dst, src in zip(self._anchor, sd["anchor"]):
            dst.copy_(src)
        for dst, src in zip(self._momentum, sd["momentum"]):
            dst.copy_(src)
        self.local_step, self.outer_step_count = sd["local_step"], sd["outer_step_count"]
