"""Global gradient-norm clipping under tensor / pipeline / data parallelism and ZeRO-1 (the reference has none; a
``torch.nn.utils.clip_grad_norm_`` on one rank's parameters computes a different norm on every rank there).

The norm is the 2-norm of the gradient of the WHOLE model, every parameter counted once:

* parameters sliced over the TENSOR group (column/row-parallel weights, the vocab-parallel table, sharded experts)
  contribute their local slice on every rank; parameters replicated over the group contribute ``1/T`` per rank;
* a table tied between the first and the last pipeline stage contributes half on each of the two stages;
* with the fused ZeRO-1 path only the slices this data-parallel rank owns hold reduced gradients (reduce-scatter), so
  each rank sums its slices and the DATA group completes the sum; otherwise every replica already holds the averaged
  gradients and the DATA group is not involved.

Partial sums of squares are all-reduced over DATA (if sliced), TENSOR and PIPELINE.  The clip coefficient
``min(1, max_norm / (norm + eps))`` is applied in place to ``p.grad`` lists, or handed to ``FusedAdam.step`` as its
``grad_scale`` (the Adam kernel multiplies the gradient while it reads it: no extra pass over the 4-byte gradients).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from pipegoose_b200.distributed.parallel_mode import ParallelMode


def _tensor_weight(p, parallel_context) -> float:
    """How much of ``|g|^2`` this rank contributes for parameter ``p``."""
    w = 1.0
    T = parallel_context.get_world_size(ParallelMode.TENSOR)
    if T > 1:
        meta = getattr(p, "parallel_metadata", None)
        sliced = bool(meta is not None and getattr(meta, "is_sliced", False))
        sharded_expert = getattr(p, "is_expert", False) and not getattr(p, "_pg_expert_replicated", False)
        if not (sliced or sharded_expert):
            w /= T
    shared = getattr(p, "_pg_pp_shared", 1)
    if shared > 1 and parallel_context.get_world_size(ParallelMode.PIPELINE) > 1:
        w /= shared
    return w


def _reduce_sum(total: torch.Tensor, parallel_context, modes) -> torch.Tensor:
    for mode in modes:
        if parallel_context.get_world_size(mode) > 1:
            group = parallel_context.get_group(mode)
            buf = total
            if dist.get_backend(group) == "nccl" and not buf.is_cuda:
                buf = buf.cuda()
            dist.all_reduce(buf, group=group)
            total = buf.to(total.device)
    return total


def _pieces(flat, params, segments) -> List[Tuple[int, int, int]]:
    """``(param index, flat start, flat end)`` for the intersections of the parameters with the owned segments."""
    out = []
    for i, p in enumerate(params):
        o, n = flat.param_range(p)
        for s, e in segments:
            a, b = max(o, s), min(o + n, e)
            if a < b:
                out.append((i, a, b))
    return out


@torch.no_grad()
def global_grad_norm(params: Iterable[torch.nn.Parameter], parallel_context, flat=None,
                     segments: Optional[List[Tuple[int, int]]] = None) -> torch.Tensor:
    """2-norm of the whole model's gradient (see the module docstring).  ``flat`` + ``segments``: read the fp32 main
    gradients of the flat buffer, restricted to the owned ``segments`` (ZeRO-1 slices); otherwise ``p.main_grad`` /
    ``p.grad`` per parameter."""
    params = [p for p in params if p.numel() > 0]
    dev = params[0].device if params else torch.device("cpu")
    sliced_over_data = False
    if flat is not None:
        whole = segments is None or (len(segments) == 1 and tuple(segments[0]) == (0, flat.numel))
        segs = [(0, flat.numel)] if whole else [tuple(s) for s in segments]
        sliced_over_data = not whole
        in_flat = [p for p in params if getattr(p, "_pg_flat_state", None) is flat]
        pieces = _pieces(flat, in_flat, segs)
        views = [flat.flat_grad[a:b] for _, a, b in pieces]
        weights = [_tensor_weight(in_flat[i], parallel_context) for i, _, _ in pieces]
    else:
        views, weights = [], []
        for p in params:
            g = getattr(p, "main_grad", None)
            g = g if g is not None else p.grad
            if g is not None:
                views.append(g.detach().reshape(-1))
                weights.append(_tensor_weight(p, parallel_context))
    if views:
        norms = torch.stack(torch._foreach_norm(views)).float()
        total = (norms * norms * torch.tensor(weights, dtype=torch.float32, device=norms.device)).sum().reshape(1)
    else:
        total = torch.zeros(1, dtype=torch.float32, device=dev)
    modes = ([ParallelMode.DATA] if sliced_over_data else []) + [ParallelMode.TENSOR, ParallelMode.PIPELINE]
    total = _reduce_sum(total, parallel_context, modes)
    return total.sqrt().reshape(())


@torch.no_grad()
def clip_grad_norm_(target, max_norm: float, parallel_context, eps: float = 1e-6) -> torch.Tensor:
    """Clip the global gradient norm to ``max_norm``; returns the norm before clipping (same value on every rank).

    ``target``: a ``DistributedOptimizer`` / ``FusedAdam`` (the coefficient is applied by the next ``step()`` through
    the fused kernel's ``grad_scale``), or an iterable of parameters / a module (gradients are scaled in place).
    Call it after ``backward()`` and before ``step()``."""
    from pipegoose_b200.optim.fused_adam import FusedAdam
    from pipegoose_b200.optim.zero.optim import DistributedOptimizer

    from pipegoose_b200.optim.diloco import DiLoCoOptimizer

    if isinstance(target, DiLoCoOptimizer):   # every worker clips its own (un-averaged) gradient: unwrap the inner optimizer
        target = target.optim
    fused = None
    if isinstance(target, DistributedOptimizer):
        if target._fused:
            if not target._zero_ready:
                target._setup_fused()
            fused = target.optim
        else:
            params = list(target._all_params)
    elif isinstance(target, FusedAdam):
        fused = target
    elif isinstance(target, torch.nn.Module):
        params = list(target.parameters())
    elif isinstance(target, torch.optim.Optimizer):
        params = [p for g in target.param_groups for p in g["params"]]
    else:
        params = list(target)

    if fused is not None:
        fused._lazy_init()
        fused._fold_autograd_grads()
        params = [p for g in fused.param_groups for p in g["params"]]
        norm = global_grad_norm(params, parallel_context, flat=fused.flat, segments=fused._segments)
    else:
        norm = global_grad_norm(params, parallel_context)
    coef = float(torch.clamp(max_norm / (norm + eps), max=1.0).item())
    if fused is not None:
        fused.pending_grad_scale = coef      # consumed (and reset) by the next FusedAdam.step
    elif coef < 1.0:
        grads = [g for g in (getattr(p, "main_grad", None) if getattr(p, "main_grad", None) is not None else p.grad
                             for p in params) if g is not None]
        if grads:
            torch._foreach_mul_(grads, coef)
        for p in params:  # a ``.grad`` materialised from main_grad in another dtype is a separate tensor
            mg = getattr(p, "main_grad", None)
            if mg is not None and p.grad is not None and p.grad.data_ptr() != mg.data_ptr():
                p.grad.mul_(coef)
    return norm
