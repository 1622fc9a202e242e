from pipegoose_b200.distributed._initializers.initialize_data import DataParallelGroupInitializer
from pipegoose_b200.distributed._initializers.initialize_expert import (
    ExpertDataParallelGroupInitializer,
    ExpertShardGroupInitializer,
)
from pipegoose_b200.distributed._initializers.initialize_pipeline import PipelineParallelGroupInitializer
from pipegoose_b200.distributed._initializers.initialize_tensor import TensorParallelGroupInitializer
from pipegoose_b200.distributed._initializers.initializer import ProcessGroupInitializer, ProcessGroupResult

__all__ = [
    "DataParallelGroupInitializer",
    "ExpertDataParallelGroupInitializer",
    "ExpertShardGroupInitializer",
    "PipelineParallelGroupInitializer",
    "TensorParallelGroupInitializer",
    "ProcessGroupInitializer",
    "ProcessGroupResult",
]
