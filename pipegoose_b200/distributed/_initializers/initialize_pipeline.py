from pipegoose_b200.distributed._initializers.initializer import ProcessGroupInitializer
from pipegoose_b200.distributed.parallel_mode import ParallelMode


class PipelineParallelGroupInitializer(ProcessGroupInitializer):
    """Creates the ``ParallelMode.PIPELINE`` groups (parity: reference distributed/_initializers/initialize_pipeline.py)."""

    parallel_mode = ParallelMode.PIPELINE
