from pipegoose_b200.nn.data_parallel.data_parallel import DataParallel

__all__ = ["DataParallel"]
