"""Pipeline exceptions (parity: reference nn/pipeline_parallel/exception.py:1-14, plus PipelineScheduleError)."""
class PipelineGradientFlowError(Exception):
    """Gradients did not flow to a stage boundary."""


class PipelineNoSavedActivationError(Exception):
    """No saved activation for the requested (microbatch, partition)."""


class PipelineNoSavedInput(Exception):
    """No saved input for the requested (microbatch, partition)."""


class PipelineScheduleError(Exception):
    """The static schedule is inconsistent (e.g. backward before forward)."""
