"""Module parallelizers: turn a leaf module into its tensor-parallel counterpart in place
(parity: reference nn/tensor_parallel/parallelizer.py:33-229).

The leaf keeps its identity (other modules may hold references to it, e.g. tied embeddings):
its class is switched to the parallel layer class and its parameters are replaced by this
rank's slice.  Sliced parameters are tagged ``param.parallel_metadata.is_sliced``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.parallel import ParallelMetadata
from pipegoose_b200.nn.tensor_parallel._utils import VocabUtility
from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
from pipegoose_b200.nn.tensor_parallel.layer_norm import LayerNorm
from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear
from pipegoose_b200.nn.tensor_parallel.parallel_mapping import TensorParallelMapping


def get_partition(data: torch.Tensor, parallel_context: ParallelContext, dim: int) -> torch.Tensor:
    rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    world = parallel_context.get_world_size(ParallelMode.TENSOR)
    assert data.size(dim) % world == 0, f"dimension {dim} of size {data.size(dim)} is not divisible by {world}"
    width = data.size(dim) // world
    return data.narrow(dim, rank * width, width).clone().contiguous()


def _mark_sliced(param: nn.Parameter, dim: int = None, full_size: int = None, is_vocab: bool = False, vocab_multiple: int = 1):
    param.parallel_metadata = ParallelMetadata(is_sliced=True, partition_dim=dim, full_size=full_size, is_vocab=is_vocab,
                                               vocab_multiple=vocab_multiple)


def _is_sliced(param) -> bool:
    meta = getattr(param, "parallel_metadata", None)
    return bool(meta is not None and meta.is_sliced)


class ModuleParallelizer(ABC):
    def __init__(self, module_name: str, module: nn.Module, model: nn.Module, parallel_context: ParallelContext):
        self.module_name = module_name
        self.module = module
        self.model = model
        self.parallel_context = parallel_context

    @staticmethod
    @abstractmethod
    def is_parallelizable(module_name: str, module: nn.Module) -> bool:
        raise NotImplementedError

    @abstractmethod
    def parallelize(self):
        raise NotImplementedError

    @abstractmethod
    def deparallelize(self):
        raise NotImplementedError


class LinearParallelizer(ModuleParallelizer):
    @staticmethod
    def is_parallelizable(module_name: str, module: nn.Module) -> bool:
        if not isinstance(module, nn.Linear):
            return False
        return TensorParallelMapping.is_column_parallel(module_name) or TensorParallelMapping.is_row_parallel(module_name)

    def parallelize(self) -> nn.Module:
        if TensorParallelMapping.is_column_parallel(self.module_name):
            return self._to_column(self.module)
        if TensorParallelMapping.is_row_parallel(self.module_name):
            return self._to_row(self.module)
        raise ValueError(f"module {self.module_name} is neither column nor row parallel")

    def deparallelize(self):
        raise NotImplementedError("gather the shards with nn.utils.save_pretrained / a fresh model instead")

    def _to_column(self, module: nn.Linear) -> nn.Module:
        ctx = self.parallel_context
        if not _is_sliced(module.weight):
            full = module.weight.shape[0]
            module.weight = nn.Parameter(get_partition(module.weight.data, ctx, dim=0), requires_grad=module.weight.requires_grad)
            _mark_sliced(module.weight, 0, full)
        if module.bias is not None and not _is_sliced(module.bias):
            full = module.bias.shape[0]
            module.bias = nn.Parameter(get_partition(module.bias.data, ctx, dim=0), requires_grad=module.bias.requires_grad)
            _mark_sliced(module.bias, 0, full)
        module.__class__ = ColumnParallelLinear
        module.gather_output = True
        module.parallel_context = ctx
        return module

    def _to_row(self, module: nn.Linear) -> nn.Module:
        ctx = self.parallel_context
        if not _is_sliced(module.weight):
            full = module.weight.shape[1]
            module.weight = nn.Parameter(get_partition(module.weight.data, ctx, dim=1), requires_grad=module.weight.requires_grad)
            _mark_sliced(module.weight, 1, full)
        module.__class__ = RowParallelLinear
        module.parallel_context = ctx
        return module


class EmbeddingParallelizer(ModuleParallelizer):
    # token-embedding names of the supported 🤗 families (Bloom / BERT / Albert, OPT / LLaMA, GPT-NeoX, GPT-2);
    # position / token-type tables and embedding subclasses with their own forward (scaled embeddings) stay as they are
    TOKEN_EMBEDDING_NAMES = ("word_embeddings", "embed_tokens", "embed_in", "wte")

    @staticmethod
    def is_parallelizable(module_name: str, module: nn.Module) -> bool:
        if "word_embeddings" in module_name:
            return isinstance(module, nn.Embedding)
        leaf = module_name.rsplit(".", 1)[-1]
        return type(module) is nn.Embedding and leaf in EmbeddingParallelizer.TOKEN_EMBEDDING_NAMES

    def parallelize(self) -> nn.Module:
        module, ctx = self.module, self.parallel_context
        world = ctx.get_world_size(ParallelMode.TENSOR)
        rank = ctx.get_local_rank(ParallelMode.TENSOR)
        if not _is_sliced(module.weight):
            old = module.weight
            weight = module.weight.data
            vocab = weight.shape[0]
            padded = (vocab + world - 1) // world * world
            if padded != vocab:  # zero-pad so that the vocabulary splits evenly
                weight = torch.cat([weight, weight.new_zeros(padded - vocab, weight.shape[1])], dim=0)
            module.weight = nn.Parameter(get_partition(weight, ctx, dim=0), requires_grad=module.weight.requires_grad)
            _mark_sliced(module.weight, 0, vocab, is_vocab=True)
            # modules that shared the table (a tied lm_head) must pick up the shard, not slice their own copy
            for other in self.model.modules():
                if other is not module and getattr(other, "weight", None) is old:
                    other._pg_tied_to_embedding = True
        else:
            padded = module.weight.shape[0] * world
        module.__class__ = ParallelEmbedding
        module.parallel_context = ctx
        module.world_size = world
        module.num_embeddings = padded
        module.vocab_start_idx, module.vocab_end_idx = VocabUtility.get_vocab_range_from_global_vocab_size(world, rank, padded)
        return module

    def deparallelize(self):
        raise NotImplementedError


class LayerNormParallelizer(ModuleParallelizer):
    @staticmethod
    def is_parallelizable(module_name: str, module: nn.Module) -> bool:
        return isinstance(module, nn.LayerNorm)

    def parallelize(self) -> nn.Module:
        module = self.module
        module.__class__ = LayerNorm
        module.parallel_context = self.parallel_context
        if not isinstance(module.normalized_shape, tuple):
            module.normalized_shape = tuple(module.normalized_shape)
        return module

    def deparallelize(self):
        self.module.__class__ = nn.LayerNorm
        return self.module


class LMHeadParallelizer(ModuleParallelizer):
    """The language-model head: column parallel with gathered output.  When its weight is tied to
    the (already sliced) input embedding, the slice is shared instead of re-sliced."""

    @staticmethod
    def is_parallelizable(module_name: str, module: nn.Module) -> bool:
        return isinstance(module, nn.Linear) and TensorParallelMapping.is_lm_head(module_name)

    def parallelize(self) -> nn.Module:
        module, ctx = self.module, self.parallel_context
        emb = self.model.get_input_embeddings() if hasattr(self.model, "get_input_embeddings") else None
        tied = emb is not None and (module.weight is emb.weight or getattr(module, "_pg_tied_to_embedding", False))
        if tied and _is_sliced(emb.weight):
            module.weight = emb.weight  # tied: share the embedding's slice
        elif not _is_sliced(module.weight):
            world = ctx.get_world_size(ParallelMode.TENSOR)
            weight = module.weight.data
            vocab = weight.shape[0]
            padded = (vocab + world - 1) // world * world
            if padded != vocab:
                weight = torch.cat([weight, weight.new_zeros(padded - vocab, weight.shape[1])], dim=0)
            module.weight = nn.Parameter(get_partition(weight, ctx, dim=0), requires_grad=module.weight.requires_grad)
            _mark_sliced(module.weight, 0, vocab, is_vocab=True)
        if module.bias is not None and not _is_sliced(module.bias):
            # heads with an output bias (BERT's ``cls.predictions.decoder``): slice it like the rows of the weight
            world = ctx.get_world_size(ParallelMode.TENSOR)
            old, bias = module.bias, module.bias.data
            full_bias = bias.shape[0]
            padded = module.weight.shape[0] * world
            if padded != bias.shape[0]:
                bias = torch.cat([bias, bias.new_zeros(padded - bias.shape[0])])
            module.bias = nn.Parameter(get_partition(bias, ctx, dim=0), requires_grad=old.requires_grad)
            _mark_sliced(module.bias, 0, full_bias, is_vocab=True)
            for other in self.model.modules():  # 🤗 keeps a second handle on the same bias (``predictions.bias``)
                if other is not module and getattr(other, "bias", None) is old:
                    other.bias = module.bias
        world = ctx.get_world_size(ParallelMode.TENSOR)
        if module.weight.shape[0] * world != module.out_features:
            module.unpadded_out_features = module.out_features   # phantom classes are cut off after the gather
        module.__class__ = ColumnParallelLinear
        module.gather_output = True
        module.parallel_context = ctx
        return module

    def deparallelize(self):
        raise NotImplementedError
