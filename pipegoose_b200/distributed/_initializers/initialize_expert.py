from pipegoose_b200.distributed._initializers.initializer import ProcessGroupInitializer
from pipegoose_b200.distributed.parallel_mode import ParallelMode


class ExpertDataParallelGroupInitializer(ProcessGroupInitializer):
    """Creates the ``ParallelMode.EXPERT_DATA`` groups (parity: reference distributed/_initializers/initialize_expert.py)."""

    parallel_mode = ParallelMode.EXPERT_DATA


class ExpertShardGroupInitializer(ProcessGroupInitializer):
    """Groups across which the experts of one MoE layer are sharded (rank sets == TENSOR groups)."""

    parallel_mode = ParallelMode.EXPERT
