"""Name-based lookup of the sub-module an ``ExpertLayer`` replaces (parity: reference
nn/expert_parallel/parallel_mapping.py:4-18; ``ExpertParallel`` itself finds the blocks structurally, this mapping is
kept for code that asks by name)."""
from pipegoose_b200.nn.parallel_mapping import ParallelInfo, ParallelMapping


class MLP(ParallelInfo):
    """Marks the feed-forward sub-module of a transformer block."""


_BLOOM_ENTRIES = [MLP(("mlp",))]


class ExpertParallelMapping(ParallelMapping):
    __MAPPING__ = {"bloom-560m": _BLOOM_ENTRIES}

    @classmethod
    def is_mlp(cls, module_name: str) -> bool:
        """``transformer.h.3.mlp`` -> True; attention / layer-norm sub-modules -> False."""
        return isinstance(cls._search(module_name), MLP)
