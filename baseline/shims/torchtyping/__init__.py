"""Dependency shim for the reference arm: `torchtyping` is not installed in this image and the
reference only uses `TensorType[...]` as a type annotation (pipegoose/nn/tensor_parallel/loss.py:6,
nn/expert_parallel/*.py).  This is a stand-in for the missing third-party package, not a change
to the reference."""


class _TensorTypeMeta(type):
    def __getitem__(cls, item):
        return cls


class TensorType(metaclass=_TensorTypeMeta):
    pass


def patch_typeguard():
    return None
