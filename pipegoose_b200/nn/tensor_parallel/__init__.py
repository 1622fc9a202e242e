from pipegoose_b200.nn.tensor_parallel.tensor_parallel import TensorParallel

__all__ = ["TensorParallel"]
