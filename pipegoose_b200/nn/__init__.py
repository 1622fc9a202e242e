"""The four parallelism wrappers.  They are resolved on first access (``from pipegoose_b200.nn import TensorParallel``
imports only the tensor-parallel stack), so tools that need one wrapper do not pay for the pipeline job runtime or the
MoE layers."""
import importlib

_WRAPPERS = {
    "DataParallel": "pipegoose_b200.nn.data_parallel.data_parallel",
    "TensorParallel": "pipegoose_b200.nn.tensor_parallel.tensor_parallel",
    "PipelineParallel": "pipegoose_b200.nn.pipeline_parallel.pipeline_parallel",
    "ExpertParallel": "pipegoose_b200.nn.expert_parallel.expert_parallel",
}
__all__ = sorted(_WRAPPERS)


def __getattr__(name: str):
    module = _WRAPPERS.get(name)
    if module is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    value = getattr(importlib.import_module(module), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + list(_WRAPPERS))
