"""Multi-process test harness (parity: reference pipegoose/testing/utils.py:16-133).

``spawn`` starts ``world_size`` local processes with a real backend (gloo on CPU, nccl on GPU
boxes); nothing is mocked.
"""
from __future__ import annotations

import os
import random
import socket
from typing import Callable

import pytest
import torch
import torch.multiprocessing as mp
from torch import nn

from pipegoose_b200.distributed.parallel_context import ParallelContext

skip_in_github_actions = pytest.mark.skipif(os.getenv("GITHUB_ACTIONS") == "true", reason="Test skipped in GitHub Actions")
skip_if_no_cuda = pytest.mark.skipif(not torch.cuda.is_available(), reason="Test requires CUDA")


def find_free_port(min_port: int = 10000, max_port: int = 32000) -> int:
    """A port nobody listens on right now.  The default range stays BELOW Linux's ephemeral range (32768-60999): the
    kernel hands ephemeral ports to outgoing connections (every gloo pair opens some), and one of those can take a port
    between this probe and the rendezvous server's ``bind`` (seen as EADDRINUSE in a busy test run)."""
    while True:
        port = random.randint(min_port, max_port)
        try:
            with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
                sock.bind(("127.0.0.1", port))
                return port
        except OSError:
            continue


def _entry(rank: int, world_size: int, port: int, func: Callable, kwargs: dict):
    if not torch.cuda.is_available() and world_size > 1:
        # CPU ranks share the host's cores: one intra-op thread each (world_size x all-cores threads only contend)
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world_size))
    func(rank=rank, world_size=world_size, port=port, **kwargs)


_FORKSERVER_PRELOAD = ["torch", "torch.distributed", "pipegoose_b200", "pipegoose_b200.nn", "pipegoose_b200.optim",
                       "pipegoose_b200.models.bloom", "pipegoose_b200.testing.utils",
                       "pipegoose_b200.nn.data_parallel.data_parallel", "pipegoose_b200.nn.tensor_parallel.tensor_parallel",
                       "pipegoose_b200.nn.pipeline_parallel.pipeline_parallel",
                       "pipegoose_b200.nn.expert_parallel.expert_parallel"]
_start_method_cache = None


def _start_method() -> str:
    """How ranks are started.  GPU hosts: ``spawn`` (a fresh interpreter per rank, the safe choice next to CUDA).
    CPU-only hosts: ``forkserver`` with the heavy modules preloaded — every rank is forked from a clean server process
    that has imported torch (and, when installed, transformers) once and has started no threads, so a rank costs
    milliseconds instead of the seconds a fresh ``import torch`` takes.  ``PIPEGOOSE_B200_START_METHOD`` overrides."""
    global _start_method_cache
    if _start_method_cache is None:
        method = os.environ.get("PIPEGOOSE_B200_START_METHOD")
        if method is None:
            method = "spawn" if torch.cuda.is_available() else "forkserver"
        if method == "forkserver":
            import importlib.util
            import multiprocessing

            preload = list(_FORKSERVER_PRELOAD)
            if importlib.util.find_spec("transformers") is not None:
                preload += ["transformers", "transformers.models.bloom.modeling_bloom"]
            try:
                multiprocessing.set_forkserver_preload(preload)
            except Exception:  # pragma: no cover - platform without forkserver
                method = "spawn"
        _start_method_cache = method
    return _start_method_cache


def spawn(func: Callable, world_size: int = 1, **kwargs):
    """Run ``func(rank, world_size, port, **kwargs)`` in ``world_size`` fresh processes."""
    pinned = kwargs.get("port") is not None
    port = kwargs.pop("port", None)
    for attempt in range(3):
        if not pinned:
            port = find_free_port()
        try:
            mp.start_processes(_entry, args=(world_size, port, func, kwargs), nprocs=world_size, join=True,
                               start_method=_start_method())
            return
        except Exception as e:
            # somebody else bound the port between the probe and the rendezvous: nothing ran yet, try another port
            if pinned or attempt == 2 or not ("EADDRINUSE" in str(e) or "address already in use" in str(e)):
                raise


def init_parallel_context(rank, world_size, port, tensor_parallel_size, pipeline_parallel_size, data_parallel_size,
                          backend: str = "gloo", host: str = "127.0.0.1", seed: int = 69) -> ParallelContext:
    return ParallelContext(
        rank=rank,
        local_rank=rank,
        world_size=world_size,
        local_world_size=world_size,
        host=host,
        port=port,
        seed=seed,
        backend=backend,
        tensor_parallel_size=tensor_parallel_size,
        pipeline_parallel_size=pipeline_parallel_size,
        data_parallel_size=data_parallel_size,
    )


N_PARTITIONS = 3      # the reference's defaults for init_pipeline_context (testing/utils.py:66-67); the number of
N_MICROBATCHES = 5    # partitions that is actually used is always the pipeline-parallel size


def init_pipeline_context(rank, world_size, port, tensor_parallel_size, pipeline_parallel_size, data_parallel_size,
                          n_partitions=None, n_microbatches=None):
    from pipegoose_b200.nn.pipeline_parallel.pipeline_context import PipelineContext
    from pipegoose_b200.nn.pipeline_parallel.scheduler import SchedulerType, get_scheduler

    parallel_context = init_parallel_context(rank, world_size, port, tensor_parallel_size, pipeline_parallel_size,
                                             data_parallel_size)
    n_partitions = n_partitions or pipeline_parallel_size
    n_microbatches = n_microbatches or N_MICROBATCHES
    scheduler = get_scheduler(SchedulerType.GPIPE)(n_microbatches, n_partitions)
    pipeline_context = PipelineContext(scheduler, parallel_context)
    return pipeline_context, parallel_context


def get_partition(data: torch.Tensor, dim: int, parallel_context: ParallelContext) -> torch.Tensor:
    from pipegoose_b200.distributed.parallel_mode import ParallelMode

    local_world_size = parallel_context.get_world_size(ParallelMode.TENSOR)
    local_rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    chunks = torch.chunk(data, chunks=local_world_size, dim=dim)
    return chunks[local_rank]


def get_microbatch(inputs, labels, parallel_context: ParallelContext, parallel_mode):
    local_rank = parallel_context.get_local_rank(parallel_mode)
    world_size = parallel_context.get_world_size(parallel_mode)
    input_chunks = torch.chunk(inputs["input_ids"], chunks=world_size)
    attention_chunks = torch.chunk(inputs["attention_mask"], chunks=world_size)
    label_chunks = torch.chunk(labels, chunks=world_size)
    return input_chunks[local_rank], attention_chunks[local_rank], label_chunks[local_rank]


def calculate_parameter_similarity(module1: nn.Module, module2: nn.Module, rtol: float = 1e-3) -> float:
    """Fraction of parameter elements that agree within ``rtol`` between two modules."""
    total, close = 0, 0
    for p1, p2 in zip(module1.parameters(), module2.parameters()):
        assert p1.size() == p2.size()
        total += p1.numel()
        close += torch.isclose(p1, p2, rtol=rtol).sum().item()
    return close / max(total, 1)


def count_model_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())
