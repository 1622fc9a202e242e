"""ZeRO-1 ``DistributedOptimizer`` (parity: reference optim/zero/optim.py:14-75).

Two execution paths behind the reference's API:

* **fused** (``optim`` is a :class:`pipegoose_b200.optim.FusedAdam`): the gradient reducer is
  switched to reduce-scatter, this rank's contiguous slice of every bucket is updated by ONE fused
  Adam launch (fp32 master/moments only for the slice), and the updated bf16 slices are
  all-gathered back into the flat parameter buffer.  Optimizer state per rank ~ 1/dp.
* **generic** (any ``torch.optim.Optimizer``): parameters are assigned to data-parallel ranks
  greedily (``OptimizerStateSharding``), the wrapped optimizer only keeps this rank's
  parameters, and after the local step each owner broadcasts its updated parameters as one flat
  tensor (the reference issues dp x groups flatten+broadcast+copy sequences).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist
from torch.optim import Optimizer

from pipegoose_b200.distributed.functional import broadcast
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.optim.base_optim import BaseDistributedOptimizer
from pipegoose_b200.optim.zero.sharding import OptimizerStateSharding
from pipegoose_b200.optim.zero.utils import copy_flatten_tensor_to_unflatten_tensors, flatten_a_list_tensor


class DistributedOptimizer(BaseDistributedOptimizer):
    def __init__(self, optim: Optimizer, parallel_context: ParallelContext):
        self.optim = optim
        self.parallel_context = parallel_context
        self.dp = parallel_context.get_world_size(ParallelMode.DATA)
        self.dp_rank = parallel_context.get_local_rank(ParallelMode.DATA)
        from pipegoose_b200.optim.fused_adam import FusedAdam

        self._fused = isinstance(optim, FusedAdam)
        self._all_params = [p for g in optim.param_groups for p in g["params"]]
        if self._fused:
            self._zero_ready = False
        else:
            self._setup_generic()

    # ------------------------------------------------------------------ generic path
    def _setup_generic(self):
        sharded = OptimizerStateSharding(self.optim.param_groups, self.parallel_context, ParallelMode.DATA).shard()
        ranks_in_group = self.parallel_context.get_ranks_in_group(ParallelMode.DATA)
        self._rank_to_param_groups: Dict[int, List[Dict]] = {rank: groups for rank, groups in zip(ranks_in_group, sharded)}
        self._local_rank_to_param_groups = sharded
        if self.dp > 1:
            self.optim.param_groups = sharded[self.dp_rank]

    def _broadcast_updated_params(self):
        replicas = self.parallel_context.get_ranks_in_group(ParallelMode.DATA)   # local data rank -> global rank
        for owner in range(self.dp):
            params = [p for g in self._local_rank_to_param_groups[owner] for p in g["params"]]
            if not params:
                continue
            flat = flatten_a_list_tensor([p.data for p in params])
            broadcast(flat, src=replicas[owner], parallel_context=self.parallel_context, parallel_mode=ParallelMode.DATA)
            if owner != self.dp_rank:
                copy_flatten_tensor_to_unflatten_tensors(flat, [p.data for p in params])

    def _params_in_shared_memory(self) -> bool:
        """Could this rank's parameters be the same storage as another rank's?  Only CPU tensors can (shared memory);
        the answer must be the same on every rank because a collective hangs on it, so it is "the model is on the CPU",
        not "this rank's tensors happen to be shared"."""
        return bool(self._all_params) and self._all_params[0].device.type == "cpu"

    def _materialize_grads_from_main(self):
        """Stock optimizers read ``p.grad``; the fused layers accumulate into ``p.main_grad``."""
        for g in self.optim.param_groups:
            for p in g["params"]:
                mg = getattr(p, "main_grad", None)
                if mg is not None and p.grad is None:
                    p.grad = mg.to(p.dtype)

    # ------------------------------------------------------------------ fused path
    def _setup_fused(self):
        optim = self.optim
        self._reducer = None
        if self.dp > 1:
            # the module's gradient reducer (installed by DataParallel) owns the flat state.  It is normally built by the
            # first forward; ``optim.zero_grad()`` BEFORE the first forward (the usual PyTorch order) builds it here.
            reducer = next((getattr(p, "_pg_dp_reducer", None) for p in self._all_params
                            if getattr(p, "_pg_dp_reducer", None) is not None), None)
            assert reducer is not None, \
                "DistributedOptimizer(FusedAdam) with dp>1 expects the module to be wrapped by DataParallel first"
            if reducer.flat is None:
                reducer.ensure_built()
            self._reducer = reducer
        flat = optim.flat
        if flat is None:
            optim.ensure_flat()
            flat = optim.flat
        if self.dp > 1:
            assert self._reducer.flat is flat, "the optimizer and the DataParallel reducer must share one flat state"
            self._reducer.mode = "reduce_scatter"
            # NVLink engine: the matrices' gradients are reduce-scattered inside the kernels that produce them
            inline = self._reducer.enable_inline_rs()
            head = getattr(self._reducer, "head", 0)
            optim.set_bucket_shards(self._reducer.zero_bucket_numel, self.dp_rank, self.dp, head=head,
                                    inline_from=head if inline else None)
        self._zero_ready = True
        provider = getattr(self, "_pending_provider", None)
        if provider is not None:
            self._pending_provider = None
            self._apply_resharded(provider)
        pending = getattr(self, "_pending_state", None)
        if pending is not None:
            self._pending_state = None
            want = [tuple(x) for x in pending[0]["segments"]]
            assert want == list(optim._segments), "the checkpointed optimizer shard does not match this layout"
            optim.load_state_dict(pending[0], *pending[1], **pending[2])

    def _all_gather_params(self):
        flat = self.optim.flat
        group = self.parallel_context.get_group(ParallelMode.DATA)
        fused = getattr(self._reducer, "_fused", None)
        head = getattr(self._reducer, "head", 0)
        if fused is not None:
            fused.all_gather_params(flat.flat_param, self._reducer.zero_bucket_numel, head)
            return
        B = self._reducer.zero_bucket_numel
        n = flat.numel
        works = []
        for start, end in ([(0, head)] if head else []) + [(s0, min(n, s0 + B)) for s0 in range(head, n, B)]:
            seg = (end - start) // self.dp
            view = flat.flat_param[start:end]
            mine = view[self.dp_rank * seg:(self.dp_rank + 1) * seg]
            if dist.get_backend(group) == "nccl":
                works.append(dist.all_gather_into_tensor(view, mine, group=group, async_op=True))
            else:
                parts = [torch.empty_like(mine) for _ in range(self.dp)]
                dist.all_gather(parts, mine.clone(), group=group)
                view.copy_(torch.cat(parts))
        for w in works:
            w.wait()

    # ------------------------------------------------------------------ optimizer API
    @property
    def defaults(self):
        return self.optim.defaults

    @property
    def param_groups(self):
        return self.optim.param_groups

    def add_param_group(self, *args, **kwargs):
        self.optim.add_param_group(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        if self._fused and not self._zero_ready and self.dp > 1:
            # the ZeRO-1 slices are laid out when the gradient reducer exists (first forward): apply the state then
            self._pending_state = (state_dict, args, kwargs)
            return
        self.optim.load_state_dict(state_dict, *args, **kwargs)

    def load_resharded_state(self, provider):
        """Elastic resume (nn/utils.py::load_training_state with another data-parallel size): ``provider(segments,
        flat_index, numel)`` builds this rank's state from all old replicas' shards once the ZeRO-1 slices of THIS job
        are laid out — now if they are, else with the first ``zero_grad()`` / ``step()``."""
        assert self._fused, "only the fused (FusedAdam) ZeRO-1 state can be re-cut for another data-parallel size"
        if not self._zero_ready and self.dp > 1:
            self._pending_provider = provider
            return
        self._apply_resharded(provider)

    def _apply_resharded(self, provider):
        from pipegoose_b200.nn.utils import _flat_index

        optim = self.optim
        optim._lazy_init()
        optim.load_state_dict(provider(list(optim._segments), _flat_index(optim), optim.flat.numel))

    def state_dict(self, *args, **kwargs):
        return self.optim.state_dict(*args, **kwargs)

    @torch.no_grad()
    def step(self, *args, **kwargs):
        if self._fused:
            if not self._zero_ready:
                self._setup_fused()
            self.optim.step(*args, **kwargs)
            if self.dp > 1:
                self._all_gather_params()
            return
        self._materialize_grads_from_main()
        if self.dp > 1 and self._params_in_shared_memory():
            # CPU parameters can live in shared memory and then are the SAME storage in every rank of this host (a module
            # handed to ``torch.multiprocessing.spawn`` as an argument — the reference's tests/optim/zero/test_optim.py
            # does that): no owner may start writing before every replica finished reading them in its backward pass
            dist.barrier(group=self.parallel_context.get_group(ParallelMode.DATA))
        self.optim.step(*args, **kwargs)
        if self.dp > 1:
            self._broadcast_updated_params()
        from pipegoose_b200.optim.fused_adam import FusedAdam

        FusedAdam.steps_taken += 1   # "the gradients were consumed" for the pipeline engine, also on ranks whose shard is empty

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Clip the norm of the whole model's gradient (every parameter counted once across the tensor / pipeline /
        data groups, optim/clip.py); call between ``backward()`` and ``step()``.  Returns the norm before clipping."""
        from pipegoose_b200.optim.clip import clip_grad_norm_

        return clip_grad_norm_(self, max_norm, self.parallel_context)

    def zero_grad(self):
        if self._fused:
            if not self._zero_ready:
                self._setup_fused()
            self.optim.zero_grad()
            return
        flat = next((p._pg_flat_state for p in self._all_params if getattr(p, "_pg_flat_state", None) is not None), None)
        held = bool(getattr(flat, "hold_grads", False))
        if flat is not None and not held:
            flat.begin_grad_window()
            flat.clears = getattr(flat, "clears", 0) + 1
        for p in self._all_params:
            p.grad = None
            if hasattr(p, "main_grad") and not held:
                # (held: a pipeline schedule produced this step's gradients inside forward and the optimizer has not
                #  consumed them yet — the fp32 main grads, which a materialised ``.grad`` may alias, must survive)
                p._mg_fresh = p.dim() >= 2
                if p.dim() < 2:
                    p.main_grad.zero_()
