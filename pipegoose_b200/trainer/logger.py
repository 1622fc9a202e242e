"""Rank-aware logger: only the chosen global rank prints (the reference's DistributedLogger is empty)."""
from __future__ import annotations

import sys
import time


class DistributedLogger:
    def __init__(self, parallel_context=None, name: str = "pipegoose_b200", rank: int = 0, stream=None):
        self.name = name
        self.parallel_context = parallel_context
        self.rank = rank
        self.stream = stream or sys.stdout
        self.records = []

    LEVELS = {"DEBUG": 10, "INFO": 20, "WARNING": 30, "ERROR": 40}

    def set_level(self, level: str = "INFO"):
        """Messages below ``level`` are recorded but not printed."""
        assert level in self.LEVELS, f"unknown level {level}"
        self.level = level
        return self

    def log(self, msg: str, level: str = "INFO"):
        self._log(level, msg)

    def _should_log(self) -> bool:
        if self.parallel_context is None:
            return True
        return self.parallel_context.get_global_rank() == self.rank

    def _log(self, level: str, msg: str):
        self.records.append((level, msg))
        if self._should_log() and self.LEVELS.get(level, 20) >= self.LEVELS[getattr(self, "level", "DEBUG")]:
            self.stream.write(f"[{time.strftime('%H:%M:%S')}] [{self.name}] [{level}] {msg}\n")
            self.stream.flush()

    def info(self, msg: str):
        self._log("INFO", msg)

    def warning(self, msg: str):
        self._log("WARNING", msg)

    def debug(self, msg: str):
        self._log("DEBUG", msg)

    def error(self, msg: str):
        self._log("ERROR", msg)


class JsonlLogger(DistributedLogger):
    """One JSON object per logged training step (``step, loss, tokens_per_s`` of this replica, ``job_tokens_per_s`` = that times the replicas, ``grad_norm, lr, tokens_seen``) appended to a
    file by the chosen global rank — what dashboards and regression checks read; text messages are recorded, not printed."""

    def __init__(self, path: str, parallel_context=None, rank: int = 0):
        import io

        super().__init__(parallel_context, name="metrics", rank=rank, stream=io.StringIO())
        self.path = path

    def log_metrics(self, metrics: dict):
        import json
        import os

        if not self._should_log():
            return
        os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
        with open(self.path, "a") as f:
            f.write(json.dumps(metrics) + "\n")
