from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode

__all__ = ["ParallelContext", "ParallelMode"]
