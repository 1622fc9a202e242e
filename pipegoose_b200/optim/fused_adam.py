"""Fused Adam/AdamW over a :class:`FlatModelState`.

fp32 master weights and moments live in flat buffers — all of them for plain training, only this
rank's slice of every gradient bucket under ZeRO-1 (``set_bucket_shards``).  A step is one kernel
launch per contiguous slice: it reads the (already averaged) fp32 gradient buffer, updates master
and moments and writes the bf16 model copy in place.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from pipegoose_b200.core.flat_state import FlatModelState
from pipegoose_b200.ops import native, use_native


class FusedAdam(torch.optim.Optimizer):
    steps_taken = 0   # process-wide count of step() calls: lets the pipeline engine tell "gradients were consumed" apart
                      # from "another accumulation micro-step" (the flat state may not exist yet during the first schedule)

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 adamw: bool = False, flat_state: Optional[FlatModelState] = None):
        params = list(params)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=adamw)
        super().__init__(params, defaults)
        for group in self.param_groups:  # (``params`` may be a list of group dicts)
            for p in group["params"]:
                p._pg_fused_optim = True  # gradients are consumed from ``main_grad``; nobody needs a ``.grad`` copy
        self.flat: Optional[FlatModelState] = flat_state
        self._segments: Optional[List[Tuple[int, int]]] = None  # [(flat_start, flat_end)] owned by this rank
        self._step = 0
        self.master = self.exp_avg = self.exp_avg_sq = None
        self.pending_grad_scale = 1.0  # set by optim.clip.clip_grad_norm_, applied once by the next step()

    def ensure_flat(self):
        if self.flat is None:
            params = [p for g in self.param_groups for p in g["params"]]
            self.flat = FlatModelState.find(params) or FlatModelState(params)
        return self.flat

    # ------------------------------------------------------------------ ZeRO-1 sharding
    def set_shard(self, start: int, end: int):
        assert self.master is None, "set_shard must be called before the first step"
        self._segments = [(start, end)]

    def set_bucket_shards(self, bucket_numel: int, rank: int, world: int, head: int = 0, inline_from: Optional[int] = None):
        """This rank owns slice ``rank`` of every gradient bucket (what reduce-scatter leaves here): the head bucket
        ``[0, head)`` (if any) and then buckets of ``bucket_numel`` elements.  ``inline_from``: gradients at and beyond
        this flat offset are reduce-scattered inside the kernels that produce them — the Adam kernel clears each owner
        slice while it reads it, so that the next step's contributions are added to zero."""
        assert self.master is None, "sharding must be fixed before the first step"
        n = self.ensure_flat().numel
        segs = []
        bounds = ([(0, head)] if head else []) + [(s, min(n, s + bucket_numel)) for s in range(head, n, bucket_numel)]
        for start, end in bounds:
            seg = (end - start) // world
            segs.append((start + rank * seg, start + (rank + 1) * seg))
        self._segments = segs
        self._inline_from = inline_from

    def _lazy_init(self):
        self.ensure_flat()
        if self._segments is None:
            self._segments = [(0, self.flat.numel)]
        if self.master is None:
            total = sum(e - s for s, e in self._segments)
            self.master = torch.empty(total, dtype=torch.float32, device=self.flat.device)
            off = 0
            for s, e in self._segments:
                self.master[off:off + e - s].copy_(self.flat.flat_param[s:e])
                off += e - s
            self.exp_avg = torch.zeros_like(self.master)
            self.exp_avg_sq = torch.zeros_like(self.master)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        self._lazy_init()
        self._fold_autograd_grads()
        grad_scale, self.pending_grad_scale = grad_scale * self.pending_grad_scale, 1.0
        self._step += 1
        plan = self._plan()
        done = set()
        if use_native(self.flat.flat_param) and len(self.param_groups) == 1:
            # ZeRO-1: the slices of equally sized buckets are equally long and equally spaced — ONE launch updates a
            # whole run of them (87 launches -> 3 for bloom-560m at dp = 2)
            g = self.param_groups[0]
            lr, (b1, b2), eps, wd = g["lr"], g["betas"], g["eps"], g["weight_decay"]
            inline = getattr(self, "_inline_from", None)
            i = 0
            while i < len(plan):
                off, s0, e0, _ = plan[i]
                n = e0 - s0
                j = i
                while (j + 1 < len(plan) and plan[j + 1][2] - plan[j + 1][1] == n
                       and (j == i or plan[j + 1][1] - plan[j][1] == plan[i + 1][1] - s0)):
                    j += 1
                if j > i and n % 4 == 0 and (plan[i + 1][1] - s0) % 4 == 0 and s0 % 4 == 0:
                    nb = j - i + 1
                    native().adam_step_strided(self.master[off:off + nb * n], self.exp_avg[off:off + nb * n],
                                               self.exp_avg_sq[off:off + nb * n], self.flat.flat_grad, self.flat.flat_param,
                                               s0, n, plan[i + 1][1] - s0, nb, lr, b1, b2, eps, wd, self._step, grad_scale,
                                               g["adamw"], inline is not None and s0 >= inline)
                    done.update(range(i, j + 1))
                i = j + 1
        for idx, (off, s, e, gi) in enumerate(plan):
            if idx in done:
                continue
            n = e - s
            g = self.param_groups[gi]
            lr, (b1, b2), eps, wd = g["lr"], g["betas"], g["eps"], g["weight_decay"]
            grad = self.flat.flat_grad[s:e]
            param = self.flat.flat_param[s:e]
            master, m, v = self.master[off:off + n], self.exp_avg[off:off + n], self.exp_avg_sq[off:off + n]
            if use_native(param):
                inline = getattr(self, "_inline_from", None)
                native().adam_step(master, m, v, grad, param, lr, b1, b2, eps, wd, self._step, grad_scale, g["adamw"],
                                   inline is not None and s >= inline)
            else:
                gr = grad.float() * grad_scale
                if wd != 0 and not g["adamw"]:
                    gr = gr + wd * master
                m.mul_(b1).add_(gr, alpha=1 - b1)
                v.mul_(b2).addcmul_(gr, gr, value=1 - b2)
                bc1, bc2 = 1 - b1 ** self._step, 1 - b2 ** self._step
                if wd != 0 and g["adamw"]:
                    master.mul_(1 - lr * wd)
                master.addcdiv_(m / bc1, (v / bc2).sqrt() + eps, value=-lr)
                param.copy_(master.to(param.dtype))
        self.flat.hold_grads = False
        self.flat.begin_grad_window()   # the gradients were consumed
        if getattr(self.flat, "inline", None) is not None:
            self.flat.inline.inline_consumed()   # ... and the owner slices of the in-kernel region cleared by the kernel
        FusedAdam.steps_taken += 1
        return loss

    def _plan(self) -> List[Tuple[int, int, int, int]]:
        """``[(offset in the optimizer state, flat start, flat end, param-group index)]``: one kernel launch each.
        One parameter group (the common case): one launch per owned segment, alignment padding included.  Several
        groups (e.g. no weight decay for biases / LayerNorms): the owned segments are cut at the group boundaries
        of the flat layout; neighbouring parameters of the same group share a launch."""
        key = (tuple(self._segments), len(self.param_groups))
        if getattr(self, "_plan_key", None) == key:
            return self._plan_cache
        plan = []
        if len(self.param_groups) == 1:
            off = 0
            for s, e in self._segments:
                plan.append((off, s, e, 0))
                off += e - s
        else:
            group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
            spans = sorted((self.flat.param_range(p)[0], sum(self.flat.param_range(p)), group_of[id(p)])
                           for p in self.flat.params if id(p) in group_of)
            off = 0
            for s, e in self._segments:
                runs = []
                for a, b, gi in spans:
                    a, b = max(a, s), min(b, e)
                    if a >= b:
                        continue
                    if runs and runs[-1][2] == gi:
                        runs[-1][1] = b       # same group: swallow the alignment gap between the two parameters
                    else:
                        runs.append([a, b, gi])
                plan.extend((off + a - s, a, b, gi) for a, b, gi in runs)
                off += e - s
        self._plan_key, self._plan_cache = key, plan
        return plan

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._plan_key = None
        for p in self.param_groups[-1]["params"]:
            p._pg_fused_optim = True

    def _fold_autograd_grads(self):
        """Gradients delivered through autograd (``p.grad``: layers without a fused wgrad, or a backward that ran
        before the flat state existed) are added to the fp32 main grads the kernels update."""
        from pipegoose_b200.ops import kernels as K

        for p in self.flat.params:
            if p.grad is not None:
                if getattr(p, "_mg_fresh", False):
                    p.main_grad.copy_(p.grad)
                    p._mg_fresh = False
                elif K.grad_rs_for(p.main_grad) is not None:
                    K.accumulate_grad(p.grad, p.main_grad, True)   # joins the in-kernel reduce-scatter
                else:
                    p.main_grad.add_(p.grad)
                p.grad = None
        self.flat.finalize_grads()

    def zero_grad(self, set_to_none: bool = True):
        self.ensure_flat().zero_grad()
        for p in self.flat.params:
            p.grad = None

    def state_dict(self):
        self._lazy_init()
        return {"step": self._step, "segments": list(self._segments), "master": self.master, "exp_avg": self.exp_avg,
                "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self._segments = [tuple(x) for x in sd["segments"]]
        self._lazy_init()
        self._step = sd["step"]
        for group, saved in zip(self.param_groups, sd.get("param_groups", [])):
            group.update({k: v for k, v in saved.items() if k != "params"})  # lr (schedulers), betas, eps, weight decay
        self.master.copy_(sd["master"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        off = 0
        for s, e in self._segments:
            self.flat.flat_param[s:e].copy_(self.master[off:off + e - s].to(self.flat.dtype))
            off += e - s
