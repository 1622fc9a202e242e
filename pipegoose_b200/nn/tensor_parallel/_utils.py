from typing import Tuple


class VocabUtility:
    """Vocabulary range owned by a tensor-parallel rank (parity: reference nn/tensor_parallel/_utils.py:4-14)."""

    @staticmethod
    def get_vocab_range_from_per_partition_vocab_size(per_partition_vocab_size: int, rank: int) -> Tuple[int, int]:
        start = rank * per_partition_vocab_size
        return start, start + per_partition_vocab_size

    @staticmethod
    def get_vocab_range_from_global_vocab_size(global_vocab_size: int, rank: int, world_size: int) -> Tuple[int, int]:
        assert global_vocab_size % world_size == 0, "the vocabulary must be padded to a multiple of the group size"
        return VocabUtility.get_vocab_range_from_per_partition_vocab_size(global_vocab_size // world_size, rank)


def is_splitable(size: int, parallel_context) -> bool:
    from pipegoose_b200.distributed.parallel_mode import ParallelMode

    return size % parallel_context.get_world_size(ParallelMode.TENSOR) == 0
