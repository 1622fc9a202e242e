from pipegoose_b200.optim.fused_adam import FusedAdam
from pipegoose_b200.optim.zero.optim import DistributedOptimizer

__all__ = ["DistributedOptimizer", "FusedAdam"]
