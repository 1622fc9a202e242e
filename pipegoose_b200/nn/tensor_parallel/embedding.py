"""Vocab-parallel embedding (parity: reference nn/tensor_parallel/embedding.py:11-42)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.tensor_parallel._functional import reduce_to_tensor_group
from pipegoose_b200.nn.tensor_parallel._utils import VocabUtility


class ParallelEmbedding(nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, parallel_context: ParallelContext):
        super().__init__()
        world = parallel_context.get_world_size(ParallelMode.TENSOR)
        rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
        assert num_embeddings % world == 0, "pad the vocabulary to a multiple of the tensor parallel size"
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.parallel_context = parallel_context
        self.world_size = world
        self.vocab_start_idx, self.vocab_end_idx = VocabUtility.get_vocab_range_from_global_vocab_size(num_embeddings, rank, world)
        self.weight = nn.Parameter(torch.empty(num_embeddings // world, embedding_dim))
        nn.init.normal_(self.weight, std=0.02)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if self.world_size > 1:
            outside = (inputs < self.vocab_start_idx) | (inputs >= self.vocab_end_idx)
            local_ids = (inputs - self.vocab_start_idx).masked_fill(outside, 0)
        else:
            outside, local_ids = None, inputs
        pad = getattr(self, "padding_idx", None)   # class-swapped nn.Embedding: the padding row still gets no gradient
        if pad is not None and self.world_size > 1:
            pad = pad - self.vocab_start_idx if self.vocab_start_idx <= pad < self.vocab_end_idx else None
        out = F.embedding(local_ids, self.weight, padding_idx=pad)
        if outside is not None:
            out = out.masked_fill(outside.unsqueeze(-1), 0.0)
        return reduce_to_tensor_group(out, self.parallel_context)
