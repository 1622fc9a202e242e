"""Vocab-parallel embedding (parity: reference nn/tensor_parallel/embedding.py:11-42): every rank of the tensor group
stores a contiguous block of the vocabulary's rows, looks up the ids that fall into its block, contributes zeros for
the rest, and the group sums the partial results.

Implemented as an ``nn.Embedding`` over the LOCAL rows, so ``padding_idx`` handling, ``extra_repr`` and the class swap
done by ``TensorParallel`` (an ``nn.Embedding`` whose weight was replaced by its slice) all come from the parent.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode
from pipegoose_b200.nn.tensor_parallel._functional import reduce_to_tensor_group
from pipegoose_b200.nn.tensor_parallel._utils import VocabUtility


class ParallelEmbedding(nn.Embedding):
    def __init__(self, num_embeddings: int, embedding_dim: int, parallel_context: ParallelContext):
        group = parallel_context.get_world_size(ParallelMode.TENSOR)
        assert num_embeddings % group == 0, "pad the vocabulary to a multiple of the tensor parallel size"
        super().__init__(num_embeddings // group, embedding_dim)
        nn.init.normal_(self.weight, std=0.02)
        self.num_embeddings = num_embeddings   # the GLOBAL vocabulary size, like the reference reports it
        self.parallel_context = parallel_context
        self.world_size = group
        self.vocab_start_idx, self.vocab_end_idx = VocabUtility.get_vocab_range_from_global_vocab_size(
            group, parallel_context.get_local_rank(ParallelMode.TENSOR), num_embeddings)

    def _local_padding_idx(self) -> Optional[int]:
        """The padding row keeps a zero gradient when it lives in this rank's block."""
        pad = getattr(self, "padding_idx", None)
        if pad is None or self.world_size == 1:
            return pad
        return pad - self.vocab_start_idx if self.vocab_start_idx <= pad < self.vocab_end_idx else None

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if self.world_size == 1:
            return F.embedding(inputs, self.weight, padding_idx=self._local_padding_idx())
        foreign = (inputs < self.vocab_start_idx) | (inputs >= self.vocab_end_idx)
        rows = (inputs - self.vocab_start_idx).masked_fill(foreign, 0)
        mine = F.embedding(rows, self.weight, padding_idx=self._local_padding_idx())
        mine = mine.masked_fill(foreign.unsqueeze(-1), 0.0)
        return reduce_to_tensor_group(mine, self.parallel_context)
