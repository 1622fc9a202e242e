from pipegoose_b200.optim.clip import clip_grad_norm_, global_grad_norm
from pipegoose_b200.optim.diloco import DiLoCoOptimizer
from pipegoose_b200.optim.fused_adam import FusedAdam
from pipegoose_b200.optim.zero.optim import DistributedOptimizer

__all__ = ["DistributedOptimizer", "DiLoCoOptimizer", "FusedAdam", "clip_grad_norm_", "global_grad_norm"]
