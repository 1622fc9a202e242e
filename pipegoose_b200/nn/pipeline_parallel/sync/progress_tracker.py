"""Accessors of the process-wide progress tracker (parity: reference nn/pipeline_parallel/sync/progress_tracker.py)."""
from pipegoose_b200.nn.pipeline_parallel.sync.handshake import get_progress_tracker, set_progress_tracker

__all__ = ["get_progress_tracker", "set_progress_tracker"]
