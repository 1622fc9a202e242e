"""Accessors of the process-wide progress tracker and the progress table of a schedule (parity: reference
nn/pipeline_parallel/sync/progress_tracker.py:6-11)."""
from typing import Dict, Tuple

from pipegoose_b200.nn.pipeline_parallel.sync.handshake import get_progress_tracker, set_progress_tracker

__all__ = ["get_progress_tracker", "set_progress_tracker", "get_progresses_from_pipeline_context"]


def get_progresses_from_pipeline_context(pipeline_context) -> Dict[int, Dict[Tuple[int, int], bool]]:
    """The table a master rank ``initiate``s a tracker with: one entry per clock cycle of the context's schedule,
    mapping every ``(microbatch_idx, partition_idx)`` task of that cycle to "not confirmed yet"."""
    return {clock: {(task.microbatch_idx, task.partition_idx): False for task in tasks}
            for clock, tasks in enumerate(pipeline_context.schedules)}
