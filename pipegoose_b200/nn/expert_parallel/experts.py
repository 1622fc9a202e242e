"""The experts of one MoE layer, sharded over the TENSOR (== EXPERT) group (parity: reference
nn/expert_parallel/experts.py:15-102).

Every rank owns ``num_local_experts`` consecutive experts.  Two token layouts are supported:

* **replicated tokens** (reference semantics, any model): each rank runs its local experts on the
  tokens routed to them and the partial outputs are summed with a differentiable all-reduce
  (the reference uses a raw in-place ``dist.all_reduce`` with no autograd, Q4);
* **token-sharded** (sequence-parallel fast path): dispatch and combine are all-to-alls
  (``distributed.functional.all_to_all`` or the fused NVLink kernels in ``ops/moe.py``).

Routing accepts either a ``[tokens]`` tensor of expert ids (the tests' ``DummyRouter``) or a
``RouterOutput`` whose ``[tokens, E]`` mask/weights are honoured: ``y = sum_k w_k * Expert_k(x)``.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Optional

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.tensor_parallel._functional import reduce_to_tensor_group


class Experts(nn.Module):
    def __init__(self, num_local_experts: int, expert: nn.Module, enable_tensor_parallel: bool,
                 parallel_context: ParallelContext):
        super().__init__()
        self.enable_tensor_parallel = enable_tensor_parallel
        self.parallel_context = parallel_context
        self.num_local_experts = num_local_experts
        # sharded: this rank holds a slice of the layer's experts and the outputs are summed over the group
        self.sharded = not enable_tensor_parallel
        self.experts = nn.ModuleList([deepcopy(expert) for _ in range(num_local_experts)])
        self._set_expert_attr(self.experts, replicated=not self.sharded)

    @staticmethod
    def _set_expert_attr(experts: nn.ModuleList, replicated: bool = False):
        # expert parameters are averaged over the EXPERT_DATA group by DataParallel
        for p in experts.parameters():
            p.is_expert = True
            p._pg_expert_replicated = replicated  # every tensor-group rank holds this expert (norms count it once)

    def _first_global_expert(self) -> int:
        if not self.sharded:
            return 0
        return self.parallel_context.get_local_rank(ParallelMode.TENSOR) * self.num_local_experts

    def forward(self, inputs: torch.Tensor, dispatch_order, *args, weights: Optional[torch.Tensor] = None,
                combine: bool = True, **kwargs):
        """``inputs``: ``[..., d]``; extra positional args that are tensors of the same token shape
        (HF Bloom passes the residual) are dispatched alongside."""
        shape = inputs.shape
        d = shape[-1]
        x = inputs.reshape(-1, d)
        n_tokens = x.shape[0]
        if dispatch_order.dim() == 1:
            ids = dispatch_order.reshape(-1)
            mask = None
        else:
            mask = dispatch_order.reshape(n_tokens, -1)
            ids = None
        w = weights.reshape(n_tokens, -1) if weights is not None else None

        extra = [a.reshape(-1, a.shape[-1]) if isinstance(a, torch.Tensor) and a.shape[:-1] == shape[:-1] else a
                 for a in args[1:]] if len(args) > 1 else []

        out = torch.zeros_like(x)
        first = self._first_global_expert()
        for local_idx, expert in enumerate(self.experts):
            e = first + local_idx
            sel = (ids == e) if ids is not None else (mask[:, e] > 0)
            rows = torch.nonzero(sel, as_tuple=False).squeeze(1)
            if rows.numel() == 0:
                continue
            xin = x.index_select(0, rows)
            extras = [a.index_select(0, rows) if isinstance(a, torch.Tensor) and a.dim() == 2 and a.shape[0] == n_tokens else a
                      for a in extra]
            y = expert(xin, *extras, **kwargs)
            if w is not None:
                y = y * w[rows, e].unsqueeze(-1).to(y.dtype)
            out = out.index_add(0, rows, y.to(out.dtype))
        if combine and self.sharded and self.parallel_context.get_world_size(ParallelMode.TENSOR) > 1:
            out = reduce_to_tensor_group(out, self.parallel_context)   # (combine=False: the caller reduce-scatters)
        return out.view(shape)

    @torch.no_grad()
    def gather_(self) -> "Experts":
        """Undo the expert sharding in place: all-gather every expert's parameters and buffers over the TENSOR group
        so that each rank holds all ``T * num_local_experts`` experts in global order (no combine afterwards)."""
        import torch.distributed as dist

        T = self.parallel_context.get_world_size(ParallelMode.TENSOR)
        if not self.sharded:
            return self
        if T > 1:
            group = self.parallel_context.get_group(ParallelMode.TENSOR)
            rank = self.parallel_context.get_local_rank(ParallelMode.TENSOR)
            El = self.num_local_experts
            full = [None] * (T * El)
            for j, expert in enumerate(self.experts):
                tensors = [t for t in list(expert.parameters()) + list(expert.buffers())]
                flat = torch.cat([t.detach().reshape(-1).float() for t in tensors]) if tensors else torch.zeros(0)
                parts = [torch.empty_like(flat) for _ in range(T)]
                dist.all_gather(parts, flat.contiguous(), group=group)
                for r in range(T):
                    clone = expert if r == rank else deepcopy(expert)
                    if r != rank:
                        off = 0
                        for t in list(clone.parameters()) + list(clone.buffers()):
                            t.copy_(parts[r][off:off + t.numel()].view_as(t).to(t.dtype))
                            off += t.numel()
                    full[r * El + j] = clone
            self.experts = nn.ModuleList(full)
            self.num_local_experts = T * El
        self.sharded = False
        self._set_expert_attr(self.experts, replicated=True)
        return self
