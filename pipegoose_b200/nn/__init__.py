from pipegoose_b200.nn.data_parallel.data_parallel import DataParallel
from pipegoose_b200.nn.tensor_parallel.tensor_parallel import TensorParallel

__all__ = ["DataParallel", "TensorParallel"]
