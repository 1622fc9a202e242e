"""Dependency shim for the reference arm: transformers 5.x removed `transformers.utils.fx`, which the
reference imports eagerly at pipegoose/nn/pipeline_parallel/partitioner.py:9 (and therefore on
`import pipegoose.nn`).  The TP/DP/ZeRO benchmark path never traces a model, so a module exposing a
`symbolic_trace` that raises if it is ever called is enough to let the unmodified reference import."""
import sys
import types


def install():
    name = "transformers.utils.fx"
    if name in sys.modules:
        return
    try:
        import transformers.utils.fx  # noqa: F401
        return
    except Exception:
        pass
    mod = types.ModuleType(name)

    def symbolic_trace(*args, **kwargs):
        raise RuntimeError("transformers.utils.fx is not available in transformers>=5 (pipeline partitioner unsupported)")

    mod.symbolic_trace = symbolic_trace
    sys.modules[name] = mod
    import transformers.utils as tu

    tu.fx = mod
