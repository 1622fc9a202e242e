"""What can go wrong inside a pipeline schedule (parity: reference nn/pipeline_parallel/exception.py:1-14; the schedule
error is new).  All share :class:`PipelineError`, and the two lookup failures say which (micro-batch, partition) slot was
asked for, because "no saved activation" without coordinates is undebuggable in a 1F1B steady state."""


class PipelineError(RuntimeError):
    pass


class _SlotError(PipelineError, KeyError):
    what = "entry"

    def __init__(self, microbatch_idx=None, partition_idx=None):
        where = "" if microbatch_idx is None else f" for micro-batch {microbatch_idx}, partition {partition_idx}"
        super().__init__(f"no saved {self.what}{where}")

    def __str__(self):
        return self.args[0]


class PipelineNoSavedActivationError(_SlotError):
    what = "activation"


class PipelineNoSavedInput(_SlotError):
    what = "input"


class PipelineGradientFlowError(PipelineError):
    """A stage boundary received no gradient: the graph between the stage's input and its output is broken."""


class PipelineInputNotRequiresGrad(PipelineError):
    """A backward job was asked for the gradient of a stage input that does not require grad."""


class PipelineScheduleError(PipelineError):
    """The static schedule is inconsistent (e.g. a backward task before its forward)."""
