"""Process-local stores of the job runtime: job queues and the activations / gradients a stage keeps
between its forward and backward jobs (parity: reference nn/pipeline_parallel/queue.py:12-124).
Everything is keyed by ``(microbatch_idx, partition_idx)`` and guarded by one lock (the reference
mutates these dicts from worker threads without synchronisation)."""
from __future__ import annotations

import threading
from queue import Queue
from typing import Any, Dict, Tuple

import torch

from pipegoose_b200.nn.pipeline_parallel.exception import PipelineNoSavedActivationError, PipelineNoSavedInput

ActivationKey = Tuple[int, int]

_LOCK = threading.RLock()
_INPUT_ACTIVATIONS: Dict[ActivationKey, Any] = {}
_SAVED_ACTIVATIONS: Dict[ActivationKey, Any] = {}
_SAVED_SCHEDULED_ACTIVATIONS: Dict[ActivationKey, Any] = {}   # last-stage outputs wrapped for the backward trigger
_SAVED_GRAD_LOSS: Dict[ActivationKey, torch.Tensor] = {}
_SAVED_METADATA_of_GRAD_LOSS: Dict[ActivationKey, Any] = {}


class JobQueue:
    """The three job queues of a process."""

    PENDING_JOBS: Queue = Queue()
    SELECTED_JOBS: Queue = Queue()
    FINISHED_JOBS: Queue = Queue()

    @classmethod
    def clear(cls):
        for q in (cls.PENDING_JOBS, cls.SELECTED_JOBS, cls.FINISHED_JOBS):
            with q.mutex:
                q.queue.clear()


class _Store:
    _data: Dict[ActivationKey, Any]

    @staticmethod
    def get_key(microbatch_idx: int, partition_idx: int) -> ActivationKey:
        return (microbatch_idx, partition_idx)

    @classmethod
    def is_saved(cls, microbatch_idx: int, partition_idx: int) -> bool:
        with _LOCK:
            return (microbatch_idx, partition_idx) in cls._data

    @classmethod
    def save_activations(cls, key: ActivationKey, data):
        with _LOCK:
            cls._data[key] = data


class SavedActivation(_Store):
    """Outputs of forward jobs, consumed by the matching backward job."""

    _data = _SAVED_ACTIVATIONS

    @classmethod
    def save_activations(cls, key: ActivationKey, data, is_by_schedule: bool = False):
        """``is_by_schedule`` (accepted and ignored by the reference, queue.py:54-56): keep the output in the store of
        the backward-trigger wrappers instead of the plain one."""
        with _LOCK:
            (_SAVED_SCHEDULED_ACTIVATIONS if is_by_schedule else cls._data)[key] = data

    @classmethod
    def get_saved_activations(cls, key: ActivationKey):
        with _LOCK:
            return cls._data.pop(key)


class InputActivations(_Store):
    """Inputs a stage received from the previous one: backward returns their ``.grad``."""

    _data = _INPUT_ACTIVATIONS

    @classmethod
    def get_saved_activations(cls, key: ActivationKey):
        with _LOCK:
            x = cls._data[key]
        if isinstance(x, torch.Tensor) and x.is_floating_point() and x.is_leaf:
            return x.requires_grad_(True)
        return x


def save_input_activations(input, microbatch_idx: int, partition_idx: int):
    InputActivations.save_activations((microbatch_idx, partition_idx), input)


def get_input_activations(microbatch_idx: int, partition_idx: int):
    try:
        return InputActivations.get_saved_activations((microbatch_idx, partition_idx))
    except KeyError:
        raise PipelineNoSavedInput(microbatch_idx, partition_idx) from None


def save_output_activations(output, microbatch_idx: int, partition_idx: int):
    SavedActivation.save_activations((microbatch_idx, partition_idx), output)


def get_output_activations(microbatch_idx: int, partition_idx: int, is_pipeline: bool = False):
    """Saved output of a forward job.  ``is_pipeline``: keep the autograd graph (the backward job will
    differentiate through it); otherwise a detached leaf is returned."""
    with _LOCK:
        try:
            out = _SAVED_ACTIVATIONS[(microbatch_idx, partition_idx)]
        except KeyError:
            raise PipelineNoSavedActivationError(microbatch_idx, partition_idx) from None
    if is_pipeline:
        return out
    return out.detach().requires_grad_(True)


def save_grad_loss(grad: torch.Tensor, microbatch_idx: int, partition_idx: int, metadata=None):
    with _LOCK:
        _SAVED_GRAD_LOSS[(microbatch_idx, partition_idx)] = grad
        _SAVED_METADATA_of_GRAD_LOSS[(microbatch_idx, partition_idx)] = metadata


def get_grad_loss(microbatch_idx: int, partition_idx: int) -> torch.Tensor:
    with _LOCK:
        return _SAVED_GRAD_LOSS.pop((microbatch_idx, partition_idx))


def clear_all():
    with _LOCK:
        for d in (_INPUT_ACTIVATIONS, _SAVED_ACTIVATIONS, _SAVED_SCHEDULED_ACTIVATIONS, _SAVED_GRAD_LOSS,
                  _SAVED_METADATA_of_GRAD_LOSS):
            d.clear()
    JobQueue.clear()
