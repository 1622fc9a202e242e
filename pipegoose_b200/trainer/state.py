"""Trainer state (parity: reference trainer/state.py; the reference Trainer never updates it, this one does)."""
from dataclasses import dataclass
from enum import Enum, auto


class TrainerStatus(Enum):
    INITIALIZING = auto()
    RUNNING = auto()
    FINISHED = auto()


class TrainerStage(Enum):
    TRAINING = auto()
    VALIDATING = auto()
    TESTING = auto()
    PREDICTING = auto()


@dataclass
class TrainerState:
    status: TrainerStatus = TrainerStatus.INITIALIZING
    stage: TrainerStage = TrainerStage.TRAINING
    epoch: int = 0
    step: int = 0
    tokens_seen: int = 0
    last_loss: float = float("nan")
    last_grad_norm: float = float("nan")
    last_eval_loss: float = float("nan")
    skipped_steps: int = 0      # optimizer steps dropped because the gradient norm was not finite
