"""Tensor parallelism wrapper (parity: reference nn/tensor_parallel/tensor_parallel.py:18-82).

``TensorParallel(module, parallel_context).parallelize()`` shards the module in place:

* ``pipegoose_b200.models`` models take the **sequence-parallel fast path**: weights are sliced,
  every block gets a :class:`TensorParallelComm`, activations between sub-layers are token
  shards and the collectives are fused into the GEMM kernels (AG->GEMM / GEMM->RS over NVLink);
  attention heads are sharded (the reference replicates attention on every rank);
* any other model (e.g. 🤗 ``BloomForCausalLM``) takes the **replicated-activation path** with the
  reference's semantics: every leaf that a parallelizer accepts becomes Column/RowParallelLinear,
  ParallelEmbedding or LayerNorm; leaves under an ``ExpertLayer`` are skipped.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.parallel import Parallel
from pipegoose_b200.nn.tensor_parallel.parallelizer import (
    EmbeddingParallelizer,
    LayerNormParallelizer,
    LinearParallelizer,
    LMHeadParallelizer,
    ModuleParallelizer,
    _mark_sliced,
    get_partition,
)


class TensorParallel(Parallel):
    PARALLELIZERS = [EmbeddingParallelizer, LinearParallelizer, LayerNormParallelizer, LMHeadParallelizer]

    def __init__(self, module: nn.Module, parallel_context: ParallelContext, sequence_parallel: Optional[bool] = None):
        """``sequence_parallel``: ``True`` forces the sequence-parallel fast path (a 🤗 ``BloomForCausalLM`` is
        converted in place to ``pipegoose_b200.models.bloom.BloomForCausalLM`` first), ``False`` forces the
        reference-style class swap, ``None`` (default) picks the fast path for ``pipegoose_b200.models`` models and for
        🤗 Bloom models that are set up for the kernels (bf16 parameters; dropout, if configured, runs through the composed
        training path) — the reference's canonical input then trains on the fused kernels without any change to the
        user's script."""
        super().__init__(module, parallel_context)
        self.sequence_parallel = sequence_parallel

    def _maybe_convert_hf(self, module: nn.Module) -> nn.Module:
        import os

        from pipegoose_b200.models.bloom import convert_hf_bloom_, hf_bloom_fast_path_blocker, is_hf_bloom

        if not is_hf_bloom(module) or self.sequence_parallel is False:
            return module
        blocker = hf_bloom_fast_path_blocker(module)
        if self.sequence_parallel is True:
            return convert_hf_bloom_(module)   # raises with the reason when it cannot
        env = os.environ.get("PIPEGOOSE_B200_HF_FAST_PATH", "auto")
        bf16 = next(module.parameters()).dtype == torch.bfloat16
        if blocker is None and env != "0" and (bf16 or env == "1"):
            return convert_hf_bloom_(module)
        return module

    @torch.no_grad()
    def parallelize(self) -> nn.Module:
        module, ctx = self.module, self.parallel_context
        module = self.module = self._maybe_convert_hf(module)   # in place: the same object, now on the fused path
        if ctx.tensor_parallel_size > 1:
            from pipegoose_b200.models.bloom import BloomForCausalLM as FastBloom

            use_sp = isinstance(module, FastBloom) if self.sequence_parallel is None else self.sequence_parallel
            if use_sp:
                assert isinstance(module, FastBloom), "the sequence-parallel path needs a pipegoose_b200.models model"
                _parallelize_fast_bloom(module, ctx)
                # partial gradients of the TP-replicated parameters are summed over the TENSOR group after every
                # backward, also when no DataParallel reducer is installed (dp == 1)
                from pipegoose_b200.core.grad_reducer import TensorPartialGradSync

                module._pg_tp_grad_sync = TensorPartialGradSync(module, ctx)
                if not hasattr(module, "no_sync") or getattr(module, "_pg_grad_reducer", None) is None:
                    module.no_sync = module._pg_tp_grad_sync.no_sync
            else:
                # remember weight tying before the embedding's parameter object is replaced by its slice
                if hasattr(module, "get_input_embeddings") and hasattr(module, "get_output_embeddings"):
                    emb, head = module.get_input_embeddings(), module.get_output_embeddings()
                    if emb is not None and head is not None and head.weight is emb.weight:
                        head._pg_tied_to_embedding = True
                for name, leaf in self._get_leaf_modules(module):
                    parallelizer = self._find_parallelizer(name, leaf)
                    if parallelizer is not None:
                        parallelizer(name, leaf, module, ctx).parallelize()
            self._save_metadata(module, ctx)
        return module

    def _get_leaf_modules(self, model: nn.Module) -> List[Tuple[str, nn.Module]]:
        """Leaves outside any ExpertLayer (experts are sharded by ExpertParallel, not sliced)."""
        from pipegoose_b200.nn.expert_parallel.layers import ExpertLayer

        expert_prefixes = [n for n, m in model.named_modules() if isinstance(m, ExpertLayer)]
        leaves = []
        for name, mod in model.named_modules():
            if any(name == p or name.startswith(p + ".") for p in expert_prefixes):
                continue
            if len(list(mod.children())) == 0:
                leaves.append((name, mod))
        return leaves

    def _find_parallelizer(self, module_name: str, module: nn.Module) -> Optional[ModuleParallelizer]:
        for parallelizer in self.PARALLELIZERS:
            if parallelizer.is_parallelizable(module_name, module):
                return parallelizer
        return None

    @torch.no_grad()
    def deparallelize(self) -> nn.Module:
        """Undo :meth:`parallelize` (unimplemented in the reference, nn/tensor_parallel/tensor_parallel.py:79-82):
        every rank all-gathers the shards and ends up with the full, unsharded module again (collective)."""
        module, ctx = self.module, self.parallel_context
        if ctx.tensor_parallel_size == 1:
            return module
        from pipegoose_b200.distributed.functional import all_gather
        from pipegoose_b200.models.bloom import BloomForCausalLM as FastBloom
        from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
        from pipegoose_b200.nn.tensor_parallel.layer_norm import LayerNorm
        from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear

        def gathered(param: nn.Parameter, dim: int, keep: Optional[int] = None) -> nn.Parameter:
            full = all_gather(param.data.contiguous(), dim=dim, parallel_context=ctx, parallel_mode=ParallelMode.TENSOR)
            if keep is not None:  # drop the zero padding that made the vocabulary divisible
                full = full.narrow(dim, 0, keep).contiguous()
            return nn.Parameter(full, requires_grad=param.requires_grad)

        if isinstance(module, FastBloom) and getattr(module, "tp", None) is not None:
            t, cfg = module.transformer, module.config
            table = gathered(t.word_embeddings.weight, 0, keep=cfg.vocab_size)
            table._pg_grad_contribs = 2
            t.word_embeddings.weight = table
            module.lm_head.weight = table
            module.vocab_start, module.tp = 0, None
            for block in t.h:
                attn, mlp = block.self_attention, block.mlp
                attn.query_key_value.weight = gathered(attn.query_key_value.weight, 0)
                attn.query_key_value.bias = gathered(attn.query_key_value.bias, 0)
                attn.dense.weight = gathered(attn.dense.weight, 1)
                attn.tp_rank = 0
                attn._slopes_cache = {}
                if hasattr(mlp, "dense_h_to_4h"):
                    mlp.dense_h_to_4h.weight = gathered(mlp.dense_h_to_4h.weight, 0)
                    mlp.dense_h_to_4h.bias = gathered(mlp.dense_h_to_4h.bias, 0)
                    mlp.dense_4h_to_h.weight = gathered(mlp.dense_4h_to_h.weight, 1)
                if hasattr(mlp, "token_comm"):
                    mlp.token_comm = None
                block.tp = None
            for p in module.parameters():
                if hasattr(p, "tp_partial_grad"):
                    del p.tp_partial_grad
            if hasattr(module, "_pg_tp_grad_sync"):
                del module._pg_tp_grad_sync
            return module

        # class-swap path (any model): gather the slices and give the leaves their torch classes back
        emb = module.get_input_embeddings() if hasattr(module, "get_input_embeddings") else None
        vocab = getattr(getattr(module, "config", None), "vocab_size", None)
        tied_table = None
        for _, leaf in self._get_leaf_modules(module):
            if isinstance(leaf, ParallelEmbedding):
                leaf.weight = gathered(leaf.weight, 0, keep=vocab)
                leaf.num_embeddings = leaf.weight.shape[0]
                for attr in ("vocab_start_idx", "vocab_end_idx", "world_size", "parallel_context"):
                    if hasattr(leaf, attr):
                        delattr(leaf, attr)
                leaf.__class__ = nn.Embedding
                if leaf is emb:
                    tied_table = leaf.weight
        for _, leaf in self._get_leaf_modules(module):
            if isinstance(leaf, ColumnParallelLinear):
                if getattr(leaf, "_pg_tied_to_embedding", False) and tied_table is not None:
                    leaf.weight = tied_table
                else:
                    is_head = leaf.weight.shape[0] * ctx.tensor_parallel_size != getattr(leaf, "out_features", -1)
                    leaf.weight = gathered(leaf.weight, 0, keep=leaf.out_features if is_head else None)
                if leaf.bias is not None:
                    leaf.bias = gathered(leaf.bias, 0)
                leaf.__class__ = nn.Linear
            elif isinstance(leaf, RowParallelLinear):
                leaf.weight = gathered(leaf.weight, 1)
                leaf.__class__ = nn.Linear
            elif isinstance(leaf, LayerNorm):
                leaf.__class__ = nn.LayerNorm
            else:
                continue
            for attr in ("gather_output", "parallel_context", "unpadded_out_features"):
                if attr in leaf.__dict__:
                    delattr(leaf, attr)
        return module


def _slice_param(param: nn.Parameter, ctx, dim: int) -> nn.Parameter:
    new = nn.Parameter(get_partition(param.data, ctx, dim=dim), requires_grad=param.requires_grad)
    _mark_sliced(new, dim, param.shape[dim])
    return new


def _parallelize_fast_bloom(model, ctx: ParallelContext):
    """Slice a pipegoose_b200 Bloom for the sequence-parallel path."""
    from pipegoose_b200.parallel.tp_comm import TensorParallelComm

    world = ctx.get_world_size(ParallelMode.TENSOR)
    rank = ctx.get_local_rank(ParallelMode.TENSOR)
    cfg = model.config
    assert cfg.n_head % world == 0, "attention heads must divide by the tensor parallel size"
    comm = TensorParallelComm(ctx)
    t = model.transformer

    # vocab-parallel (tied) embedding / lm_head, zero-padded to a multiple of the group size
    table = t.word_embeddings.weight.data
    vocab = table.shape[0]
    padded = (vocab + world * 8 - 1) // (world * 8) * (world * 8)
    if padded != vocab:
        table = torch.cat([table, table.new_zeros(padded - vocab, table.shape[1])], dim=0)
    sliced = nn.Parameter(get_partition(table, ctx, dim=0))
    _mark_sliced(sliced, 0, vocab, is_vocab=True, vocab_multiple=8)
    sliced._pg_grad_contribs = 2  # lm_head wgrad + embedding backward
    t.word_embeddings.weight = sliced
    model.lm_head.weight = sliced
    model.vocab_start = rank * (padded // world)
    model.tp = comm

    def partial(p):  # gradient is a partial sum over this rank's token shard
        p.tp_partial_grad = True

    for ln in (getattr(t, "word_embeddings_layernorm", None), t.ln_f):
        if ln is not None:
            partial(ln.weight), partial(ln.bias)
    if hasattr(t, "position_embeddings"):  # replicated; every rank sees the positions of its token shard only
        partial(t.position_embeddings.weight)
    for block in t.h:
        attn, mlp = block.self_attention, block.mlp
        attn.query_key_value.weight = _slice_param(attn.query_key_value.weight, ctx, 0)  # whole heads
        attn.query_key_value.bias = _slice_param(attn.query_key_value.bias, ctx, 0)
        attn.dense.weight = _slice_param(attn.dense.weight, ctx, 1)
        attn.tp_rank = rank
        partial(attn.dense.bias)
        if hasattr(mlp, "dense_h_to_4h"):
            mlp.dense_h_to_4h.weight = _slice_param(mlp.dense_h_to_4h.weight, ctx, 0)
            mlp.dense_h_to_4h.bias = _slice_param(mlp.dense_h_to_4h.bias, ctx, 0)
            mlp.dense_4h_to_h.weight = _slice_param(mlp.dense_4h_to_h.weight, ctx, 1)
            partial(mlp.dense_4h_to_h.bias)
        for ln in (block.input_layernorm, block.post_attention_layernorm):
            partial(ln.weight), partial(ln.bias)
        from pipegoose_b200.nn.expert_parallel.layers import ExpertLayer

        if isinstance(mlp, ExpertLayer):
            # token-sharded activations: the layer exchanges tokens itself (all-gather / reduce-scatter around its local
            # experts); the router's gradients are partial sums over the group
            mlp.token_comm = comm
            for p in mlp.router.parameters():
                partial(p)
        block.tp = comm
    hooks = getattr(model, "_pg_after_move_hooks", [])
    hooks.append(lambda m: comm.enable_fused())
    model._pg_after_move_hooks = hooks
