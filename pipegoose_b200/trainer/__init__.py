from pipegoose_b200.trainer.callback import Callback
from pipegoose_b200.trainer.logger import DistributedLogger, JsonlLogger
from pipegoose_b200.trainer.state import TrainerStage, TrainerState, TrainerStatus
from pipegoose_b200.trainer.trainer import Trainer

__all__ = ["Trainer", "Callback", "DistributedLogger", "JsonlLogger", "TrainerState", "TrainerStatus", "TrainerStage"]
