from enum import Enum


class ParallelMode(Enum):
    """Kinds of process group a rank belongs to (parity: reference distributed/parallel_mode.py:4-12).

    ``EXPERT`` is an addition: the group across which *different* experts are sharded (same rank
    sets as ``TENSOR``).  ``EXPERT_DATA`` is the group whose members hold the *same* experts and
    therefore average expert gradients (see DESIGN.md "expert-data semantics").
    """

    GLOBAL = "global"

    TENSOR = "tensor"
    PIPELINE = "pipeline"
    DATA = "data"

    EXPERT = "expert_shard"
    EXPERT_DATA = "expert"
