from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext
from pipegoose_b200.nn.expert_parallel.expert_parallel import ExpertParallel
from pipegoose_b200.nn.expert_parallel.layers import ExpertLayer
from pipegoose_b200.nn.expert_parallel.loss import ExpertLoss
from pipegoose_b200.nn.expert_parallel.routers import RouterOutput, SwitchNoisePolicy, Top1Router, Top2Router

__all__ = ["ExpertParallel", "ExpertLoss", "ExpertLayer", "ExpertContext", "Top1Router", "Top2Router",
           "SwitchNoisePolicy", "RouterOutput"]
