"""Tracing helpers (the reference has no runtime tracing at all): NVTX ranges that show up in
Nsight timelines, and a device-side timer whose readings follow the bench contract — CUDA events on the
launching stream, synchronised on both sides, max over the ranks of a group."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, List

import torch


@contextmanager
def nvtx_range(name: str):
    """``with nvtx_range("fwd"): ...`` — a no-op on machines without CUDA."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class DeviceTimer:
    """Accumulates device time per label.

        timer = DeviceTimer(parallel_context)          # group = GLOBAL by default
        with timer("step"): train_step()
        timer.summary()  ->  {"step": {"count": n, "ms_max_over_ranks": t}}

    Regions are measured with CUDA events (``time.perf_counter`` on CPU-only runs); ``summary`` synchronises
    once and reduces with MAX over the group, so a multi-GPU number is never a wall-clock number of one rank."""

    def __init__(self, parallel_context=None, parallel_mode=None):
        self.ctx = parallel_context
        self.mode = parallel_mode
        self._events: Dict[str, List] = {}
        self._cpu: Dict[str, float] = {}
        self._counts: Dict[str, int] = {}

    @contextmanager
    def __call__(self, label: str):
        self._counts[label] = self._counts.get(label, 0) + 1
        if torch.cuda.is_available():
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with nvtx_range(label):
                s.record()
                yield
                e.record()
            self._events.setdefault(label, []).append((s, e))
        else:
            import time

            t0 = time.perf_counter()
            yield
            self._cpu[label] = self._cpu.get(label, 0.0) + (time.perf_counter() - t0) * 1e3

    def summary(self, reset: bool = True) -> Dict[str, Dict[str, float]]:
        import torch.distributed as dist

        out = {}
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        labels = sorted(set(self._events) | set(self._cpu))
        for label in labels:
            ms = self._cpu.get(label, 0.0) + sum(s.elapsed_time(e) for s, e in self._events.get(label, []))
            if self.ctx is not None and dist.is_initialized():
                from pipegoose_b200.distributed.parallel_mode import ParallelMode

                mode = self.mode or ParallelMode.GLOBAL
                if self.ctx.get_world_size(mode) > 1:
                    dev = self.ctx.device if torch.cuda.is_available() and dist.get_backend(self.ctx.get_group(mode)) == "nccl" else "cpu"
                    t = torch.tensor([ms], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.ctx.get_group(mode))
                    ms = float(t.item())
            out[label] = {"count": self._counts.get(label, 0), "ms_max_over_ranks": ms}
        if reset:
            self._events.clear(), self._cpu.clear(), self._counts.clear()
        return out
