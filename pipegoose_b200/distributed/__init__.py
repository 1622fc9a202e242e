"""Distributed runtime: process-group bookkeeping (``ParallelContext`` / ``ParallelMode``), the arithmetic rank
``Topology``, functional collectives and the NVLink symmetric workspace."""
from pipegoose_b200.distributed.parallel_mode import ParallelMode  # noqa: F401  (first: parallel_context imports it)
from pipegoose_b200.distributed.parallel_context import ParallelContext  # noqa: F401
from pipegoose_b200.distributed.topology import Topology  # noqa: F401

__all__ = ["ParallelContext", "ParallelMode", "Topology"]
