"""Hooks of the progress tracker (parity: reference nn/pipeline_parallel/sync/callback.py:4-8).  A callback is fired on
every rank each time a clock cycle of the published schedule completes; lower ``order`` runs first."""
from typing import Dict


class Callback:
    order = 0

    @property
    def name(self) -> str:
        return type(self).__name__

    def after_new_clock_cycle(self, progress: Dict[int, Dict[object, bool]], clock_idx: int):
        """``progress``: ``{clock: {task: done}}`` after the update; ``clock_idx``: the cycle that starts now."""
