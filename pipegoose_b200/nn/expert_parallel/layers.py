"""``ExpertLayer``: router + sharded experts (parity: reference nn/expert_parallel/layers.py:11-48).

Drop-in replacement for a transformer block's MLP.  With HF Bloom's calling convention
``mlp(hidden_states, residual)`` the residual is kept out of the experts:
``y = residual + sum_k w_k * Expert_k(x)`` (tokens dropped by the capacity limit pass through on
the residual path only).
"""
from __future__ import annotations

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext
from pipegoose_b200.nn.expert_parallel.experts import Experts
from pipegoose_b200.nn.expert_parallel.routers import RouterOutput
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.expert_parallel.utils import get_num_local_experts
from pipegoose_b200.nn.tensor_parallel._functional import broadcast_to_tensor_group


class ExpertLayer(nn.Module):
    def __init__(self, num_experts: int, expert: nn.Module, router: nn.Module, enable_tensor_parallel: bool,
                 parallel_context: ParallelContext):
        super().__init__()
        self.num_experts = num_experts
        self.router = router
        self.parallel_context = parallel_context
        if enable_tensor_parallel:
            self.num_local_experts = num_experts
        else:
            self.num_local_experts = get_num_local_experts(num_experts, parallel_context)
        self._experts = Experts(self.num_local_experts, expert, enable_tensor_parallel, parallel_context)

    @property
    def experts(self) -> nn.ModuleList:
        return self._experts.experts

    @torch.no_grad()
    def gather_experts_(self) -> "ExpertLayer":
        """Every rank ends up with all ``num_experts`` experts (used by ``ExpertParallel.deparallelize``)."""
        self._experts.gather_()
        self.num_local_experts = self._experts.num_local_experts
        return self

    def forward(self, *args, **kwargs) -> torch.Tensor:
        inputs = args[0]
        comm = getattr(self, "token_comm", None)   # set by the sequence-parallel TensorParallel path: tokens are sharded
        group_size = comm.size if comm is not None else 1
        sharded_tokens = comm is not None and group_size > 1
        # sharded experts + sharded tokens: every rank needs ALL tokens to feed the experts it owns.  Megatron-SP shape:
        # all-gather the tokens, run the local experts on whatever is routed to them, reduce-scatter the partial sums.
        exchange = sharded_tokens and self._experts.sharded
        local_shape = inputs.shape
        if exchange:
            inputs = comm.gather_rows(inputs.reshape(-1, local_shape[-1]))
        routed = self.router(inputs)
        ctx = ExpertContext.get_instance()
        if isinstance(routed, RouterOutput):
            aux, z = routed.aux_loss, routed.z_loss
            if sharded_tokens:
                # every rank of the group adds these terms to ITS loss while the router parameters' gradients are summed
                # over the group: keep the value, give each rank 1/T of the gradient
                aux, z = _scale_grad(aux, 1.0 / group_size), _scale_grad(z, 1.0 / group_size)
            ctx.push_aux_loss(aux)
            ctx.push_z_loss(z)
            order, weights = routed.dispatching_order, routed.weight
        else:  # bare expert ids, e.g. a test router
            order, weights = routed, None
        residual = None
        rest = list(args[1:])
        if rest and isinstance(rest[0], torch.Tensor) and rest[0].shape == local_shape:
            # HF Bloom: mlp(layernorm_output, residual) -> keep the residual outside the experts
            residual = rest[0]
            rest[0] = torch.zeros_like(inputs)
        expert_inputs = inputs
        if self._experts.sharded and not exchange and self.parallel_context.get_world_size(ParallelMode.TENSOR) > 1:
            # replicated tokens (class-swap TP, the reference's layout): every rank adds only its own experts' terms to
            # the all-reduced output, so the gradients that flow back into the tokens and into the routing weights are
            # partial sums — the conjugate op (identity forward, all-reduce backward) completes them on every rank
            expert_inputs = broadcast_to_tensor_group(inputs, self.parallel_context)
            if weights is not None:
                weights = broadcast_to_tensor_group(weights, self.parallel_context)
        out = self._experts(expert_inputs, order, expert_inputs, *rest, weights=weights, combine=not exchange, **kwargs)
        if expert_inputs is not inputs or exchange:
            # The communication ops around the local experts have collective BACKWARD passes.  A rank whose experts got
            # no token this step would otherwise have no autograd path through them and skip those collectives while
            # its peers wait: tie the output to the exchanged tensors so that every rank runs the same backward graph.
            anchor = expert_inputs.reshape(-1)[0] * 0.0
            if isinstance(weights, torch.Tensor) and weights.requires_grad:
                anchor = anchor + weights.reshape(-1)[0] * 0.0
            out = out + anchor.to(out.dtype)
        if exchange:
            out = comm.scatter_rows(out.reshape(-1, local_shape[-1])).view(local_shape)
        if residual is not None:
            out = out + residual
        return out


def _scale_grad(t, factor: float):
    """Same value, gradient multiplied by ``factor``."""
    if not isinstance(t, torch.Tensor) or not t.requires_grad:
        return t
    return t.detach() + (t - t.detach()) * factor
