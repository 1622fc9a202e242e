from enum import Enum, auto


class JobType(Enum):
    FORWARD = auto()
    BACKWARD = auto()
