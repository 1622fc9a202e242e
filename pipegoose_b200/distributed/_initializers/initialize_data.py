from pipegoose_b200.distributed._initializers.initializer import ProcessGroupInitializer
from pipegoose_b200.distributed.parallel_mode import ParallelMode


class DataParallelGroupInitializer(ProcessGroupInitializer):
    """Creates the ``ParallelMode.DATA`` groups (parity: reference distributed/_initializers/initialize_data.py)."""

    parallel_mode = ParallelMode.DATA
