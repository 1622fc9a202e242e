from pipegoose_b200.nn.data_parallel.data_parallel import DataParallel
from pipegoose_b200.nn.expert_parallel.expert_parallel import ExpertParallel
from pipegoose_b200.nn.pipeline_parallel.pipeline_parallel import PipelineParallel
from pipegoose_b200.nn.tensor_parallel.tensor_parallel import TensorParallel

__all__ = ["DataParallel", "TensorParallel", "PipelineParallel", "ExpertParallel"]
