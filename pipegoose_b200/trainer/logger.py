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
