"""Vocabulary ranges of the tensor-parallel ranks (parity: reference nn/tensor_parallel/_utils.py:4-14 — same method names
and argument order: ``(partition_size, rank)`` and ``(world_size, rank, vocab_size)``)."""
from typing import Tuple


class VocabUtility:
    @staticmethod
    def get_vocab_range_idx_from_partition_size(partition_size: int, rank: int) -> Tuple[int, int]:
        """Rows ``[rank * partition_size, (rank + 1) * partition_size)`` of the (padded) vocabulary."""
        first = rank * partition_size
        return first, first + partition_size

    @staticmethod
    def get_vocab_range_from_global_vocab_size(world_size: int, rank: int, vocab_size: int) -> Tuple[int, int]:
        assert vocab_size % world_size == 0, "the vocabulary must be padded to a multiple of the group size"
        return VocabUtility.get_vocab_range_idx_from_partition_size(vocab_size // world_size, rank)

    # Megatron's spelling, kept as an alias
    get_vocab_range_from_per_partition_vocab_size = get_vocab_range_idx_from_partition_size


def is_splitable(size: int, parallel_context) -> bool:
    from pipegoose_b200.distributed.parallel_mode import ParallelMode

    return size % parallel_context.get_world_size(ParallelMode.TENSOR) == 0
