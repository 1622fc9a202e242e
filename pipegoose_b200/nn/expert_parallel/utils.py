"""Expert-parallel helpers (parity: reference nn/expert_parallel/utils.py:5-8)."""
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def get_num_local_experts(num_experts: int, parallel_context: ParallelContext) -> int:
    """Experts per rank: the experts of a layer are sharded over the TENSOR (== EXPERT) group."""
    tensor_parallel_size = parallel_context.get_world_size(ParallelMode.TENSOR)
    assert num_experts % tensor_parallel_size == 0, "num_experts must be divisible by the tensor parallel size"
    return num_experts // tensor_parallel_size
