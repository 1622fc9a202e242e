from pipegoose_b200.distributed._initializers.initializer import ProcessGroupInitializer
from pipegoose_b200.distributed.parallel_mode import ParallelMode


class TensorParallelGroupInitializer(ProcessGroupInitializer):
    """Creates the ``ParallelMode.TENSOR`` groups (parity: reference distributed/_initializers/initialize_tensor.py)."""

    parallel_mode = ParallelMode.TENSOR
