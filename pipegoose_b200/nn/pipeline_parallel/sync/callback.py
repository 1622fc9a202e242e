"""Callbacks of the progress tracker (parity: reference nn/pipeline_parallel/sync/callback.py:4-8)."""


class Callback:
    order = 0

    def after_new_clock_cycle(self, progress, clock_idx):
        pass
