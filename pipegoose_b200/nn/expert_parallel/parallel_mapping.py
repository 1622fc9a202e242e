from pipegoose_b200.nn.parallel_mapping import ParallelInfo, ParallelMapping


class MLP(ParallelInfo):
    pass


class ExpertParallelMapping(ParallelMapping):
    """Which sub-module of a block is replaced by an expert layer (parity: reference
    nn/expert_parallel/parallel_mapping.py:4-18)."""

    __MAPPING__ = {
        "bloom-560m": [MLP(("mlp",))],
    }

    @staticmethod
    def is_mlp(module_name: str) -> bool:
        return isinstance(ExpertParallelMapping._search(module_name), MLP)
