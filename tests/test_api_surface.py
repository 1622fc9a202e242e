"""API-surface parity: every import path, class, function and attribute that the reference's README, examples and tests
touch resolves in ``pipegoose_b200`` with the same call shape.  Pure imports / signature inspection —
behaviour is covered by the other test modules."""
import importlib
import inspect

import pytest

SURFACE = {
    "pipegoose_b200.distributed": ["ParallelContext", "ParallelMode"],
    "pipegoose_b200.distributed.parallel_context": ["ParallelContext"],
    "pipegoose_b200.distributed.parallel_mode": ["ParallelMode"],
    "pipegoose_b200.distributed.functional": [
        "scatter", "reduce", "broadcast", "all_gather", "all_reduce", "reduce_scatter", "send", "recv", "barrier"],
    "pipegoose_b200.nn": ["DataParallel", "TensorParallel", "PipelineParallel", "ExpertParallel"],
    "pipegoose_b200.nn.tensor_parallel.linear": ["ColumnParallelLinear", "RowParallelLinear"],
    "pipegoose_b200.nn.tensor_parallel.embedding": ["ParallelEmbedding"],
    "pipegoose_b200.nn.tensor_parallel.layer_norm": ["LayerNorm"],
    "pipegoose_b200.nn.tensor_parallel.loss": ["VocabParallelCrossEntropy"],
    "pipegoose_b200.nn.tensor_parallel.parallelizer": [
        "EmbeddingParallelizer", "LinearParallelizer", "LayerNormParallelizer", "LMHeadParallelizer"],
    "pipegoose_b200.nn.tensor_parallel.parallel_mapping": ["TensorParallelMapping"],
    "pipegoose_b200.nn.tensor_parallel._utils": ["VocabUtility"],
    "pipegoose_b200.nn.expert_parallel": [
        "ExpertParallel", "ExpertLoss", "Top1Router", "Top2Router", "SwitchNoisePolicy"],
    "pipegoose_b200.nn.expert_parallel.routers": ["RouterOutput", "Top1Router", "Top2Router", "SwitchNoisePolicy"],
    "pipegoose_b200.nn.expert_parallel.layers": ["ExpertLayer"],
    "pipegoose_b200.nn.expert_parallel.expert_context": ["ExpertContext"],
    "pipegoose_b200.nn.expert_parallel.utils": ["get_num_local_experts"],
    "pipegoose_b200.nn.pipeline_parallel": ["PipelineParallel"],
    "pipegoose_b200.nn.pipeline_parallel.scheduler": ["GPipeScheduler", "SchedulerType", "get_scheduler"],
    "pipegoose_b200.nn.pipeline_parallel.partitioner": ["UniformPartitioner", "BasePartitioner", "PartitionPolicy", "INPUT_NAMES",
                                                        "get_model_partition"],
    "pipegoose_b200.nn.pipeline_parallel.pipeline_engine": ["PipelineEngine", "Schedule"],
    "pipegoose_b200.nn.pipeline_parallel._job.creator": [
        "create_job", "schedule_backward_job", "schedule_backward_execution", "ScheduleBackwardJobCallback", "JobCreator"],
    "pipegoose_b200.nn.pipeline_parallel.microbatch": ["split"],
    "pipegoose_b200.nn.pipeline_parallel.pipeline_context": ["PipelineContext"],
    "pipegoose_b200.nn.pipeline_parallel._utils": ["get_partition_idx", "is_last_stage"],
    "pipegoose_b200.optim": ["DistributedOptimizer"],
    "pipegoose_b200.optim.zero.optim": ["DistributedOptimizer"],
    "pipegoose_b200.optim.zero.sharding": ["OptimizerStateSharding"],
    "pipegoose_b200.optim.zero.utils": ["flatten_a_list_tensor", "copy_flatten_tensor_to_unflatten_tensors"],
    "pipegoose_b200.nn.utils": ["save_pretrained", "from_pretrained"],
    "pipegoose_b200.constants": ["SEED", "CHECKPOINT_WEIGHTS_NAME", "CHECKPOINT_PATH_NAME", "BUCKET_SIZE_MB"],
    "pipegoose_b200.core.bucket.bucket": ["Bucket"],
    "pipegoose_b200.core.bucket.dist": ["BucketDistributor"],
    "pipegoose_b200.core.bucket.utils": ["mb_size_to_num_elements"],
    "pipegoose_b200.core.bucket.exception": ["BucketFullError", "BucketClosedError"],
    "pipegoose_b200.testing.utils": [
        "spawn", "init_parallel_context", "find_free_port", "skip_if_no_cuda", "skip_in_github_actions",
        "get_partition", "calculate_parameter_similarity", "count_model_parameters", "init_pipeline_context", "get_microbatch",
        "N_PARTITIONS", "N_MICROBATCHES"],
    "pipegoose_b200.trainer": ["Trainer", "Callback", "DistributedLogger", "TrainerState"],
}


@pytest.mark.parametrize("module_name", sorted(SURFACE))
def test_module_exports(module_name):
    mod = importlib.import_module(module_name)
    missing = [n for n in SURFACE[module_name] if not hasattr(mod, n)]
    assert not missing, f"{module_name} lacks {missing}"


def test_parallel_context_members():
    from pipegoose_b200.distributed import ParallelContext

    sig = inspect.signature(ParallelContext.from_torch)
    assert list(sig.parameters)[:3] == ["tensor_parallel_size", "pipeline_parallel_size", "data_parallel_size"]
    assert sig.parameters["backend"].default == "gloo"
    assert "seed" in sig.parameters
    for name in (
        "get_context", "get_global_rank", "get_local_rank", "get_world_size", "get_group", "get_ranks_in_group",
        "get_global_rank_from_local_rank", "get_next_global_rank", "get_prev_global_rank", "get_next_local_rank",
        "get_prev_local_rank", "is_first_rank", "is_last_rank", "is_initialized", "ranks2device", "get_worker_name",
        "set_device", "set_seed", "destroy", "add_local_rank", "add_global_rank", "add_world_size", "add_group",
        "add_ranks_in_group",
    ):
        assert callable(getattr(ParallelContext, name, None)), name


def test_parallel_mode_members():
    from pipegoose_b200.distributed import ParallelMode

    for name in ("GLOBAL", "TENSOR", "PIPELINE", "DATA", "EXPERT_DATA"):
        assert hasattr(ParallelMode, name), name


def _params(obj):
    return list(inspect.signature(obj).parameters)


def test_layer_constructor_shapes():
    from pipegoose_b200.nn.expert_parallel import ExpertLoss, SwitchNoisePolicy, Top1Router
    from pipegoose_b200.nn.expert_parallel.layers import ExpertLayer
    from pipegoose_b200.nn.tensor_parallel.embedding import ParallelEmbedding
    from pipegoose_b200.nn.tensor_parallel.layer_norm import LayerNorm
    from pipegoose_b200.nn.tensor_parallel.linear import ColumnParallelLinear, RowParallelLinear

    assert _params(ColumnParallelLinear)[:5] == ["in_features", "out_features", "bias", "gather_output", "parallel_context"]
    assert _params(RowParallelLinear)[:4] == ["in_features", "out_features", "bias", "parallel_context"]
    assert _params(ParallelEmbedding)[:3] == ["num_embeddings", "embedding_dim", "parallel_context"]
    assert _params(LayerNorm)[:4] == ["normalized_shape", "eps", "bias", "parallel_context"]
    assert _params(ExpertLoss)[:3] == ["loss_func", "aux_weight", "z_weight"]
    assert _params(Top1Router)[:6] == ["noise_policy", "num_experts", "d_model", "expert_capacity", "alpha", "eps"]
    assert _params(SwitchNoisePolicy)[:1] == ["eps"]
    assert _params(ExpertLayer)[:5] == ["num_experts", "expert", "router", "enable_tensor_parallel", "parallel_context"]


def test_wrapper_and_parallelizer_shapes():
    from pipegoose_b200.nn import DataParallel, ExpertParallel, PipelineParallel, TensorParallel
    from pipegoose_b200.nn.pipeline_parallel.scheduler import GPipeScheduler
    from pipegoose_b200.nn.tensor_parallel.parallel_mapping import TensorParallelMapping
    from pipegoose_b200.nn.tensor_parallel.parallelizer import LinearParallelizer
    from pipegoose_b200.nn.utils import from_pretrained, save_pretrained
    from pipegoose_b200.optim import DistributedOptimizer

    assert _params(TensorParallel)[:2] == ["module", "parallel_context"]
    assert _params(DataParallel)[:2] == ["module", "parallel_context"]
    assert _params(PipelineParallel)[:3] == ["module", "num_microbatches", "parallel_context"]
    assert _params(ExpertParallel)[:2] == ["module", "num_experts"]
    assert "parallel_context" in _params(ExpertParallel)
    for cls in (TensorParallel, DataParallel, PipelineParallel, ExpertParallel):
        assert callable(getattr(cls, "parallelize"))
        assert callable(getattr(cls, "deparallelize"))
    assert _params(LinearParallelizer)[:4] == ["module_name", "module", "model", "parallel_context"]
    assert callable(LinearParallelizer.is_parallelizable)
    for fn in ("is_column_parallel", "is_row_parallel", "is_lm_head"):
        assert callable(getattr(TensorParallelMapping, fn))
    assert _params(GPipeScheduler)[:2] == ["n_microbatches", "n_partitions"]
    for fn in ("get_schedules", "get_forward_schedules", "get_backward_schedules"):
        assert callable(getattr(GPipeScheduler, fn))
    for prop in ("total_clock_cycles", "total_forward_clock_cycles", "total_backward_clock_cycles"):
        assert hasattr(GPipeScheduler, prop)
    assert _params(DistributedOptimizer)[:2] == ["optim", "parallel_context"]
    for member in ("defaults", "param_groups", "add_param_group", "load_state_dict", "state_dict", "step", "zero_grad"):
        assert hasattr(DistributedOptimizer, member), member
    assert {"module", "ckp_name", "ckp_path", "parallel_context"} <= set(_params(save_pretrained))
    assert _params(from_pretrained)[:3] == ["module", "ckp_path", "parallel_context"]


def test_round2_public_surface():
    """Names added in round 2 that scripts and docs rely on."""
    import inspect

    from pipegoose_b200.distributed import symmetric
    from pipegoose_b200.models import bloom
    from pipegoose_b200.nn import TensorParallel
    from pipegoose_b200.nn.pipeline_parallel.partitioner import UniformPartitioner
    from pipegoose_b200.nn.utils import capture_rng_state, from_pretrained, restore_rng_state
    from pipegoose_b200.ops import kernels as K

    assert callable(bloom.convert_hf_bloom_) and callable(bloom.is_hf_bloom) and callable(bloom.hf_bloom_fast_path_blocker)
    assert "sequence_parallel" in inspect.signature(TensorParallel.__init__).parameters
    assert callable(UniformPartitioner.register_family)
    assert "strict" in inspect.signature(from_pretrained).parameters
    assert set(capture_rng_state()) >= {"torch", "cuda", "python", "numpy"} and callable(restore_rng_state)
    assert callable(symmetric.exchange_fds) and hasattr(symmetric.SymmetricWorkspace, "mc_data_ptr")
    for name in ("register_grad_rs", "grad_rs_for", "ce_partials_buffer", "ce_stats_from_partials"):
        assert callable(getattr(K, name)), name


def test_operations_layer_public_surface():
    """Names of the operations layer (docs/OPERATIONS.md) that scripts and docs rely on."""
    import inspect

    from pipegoose_b200.models.bloom import BloomConfig, BloomForCausalLM
    from pipegoose_b200.nn import checkpoint_convert as cc
    from pipegoose_b200.nn import utils as nu
    from pipegoose_b200.optim import DistributedOptimizer
    from pipegoose_b200.partitioning import planner
    from pipegoose_b200.trainer import Callback, Trainer
    from pipegoose_b200.utils import data, watchdog

    for name in ("consolidate_checkpoint", "reshard_checkpoint", "shard_state_dict", "shard_layout", "write_layout", "main"):
        assert callable(getattr(cc, name)), name
    assert callable(nu.reshard_fused_state) and callable(DistributedOptimizer.load_resharded_state)
    for name in ("data_parallel_sampler", "build_dataloader", "TokenFileDataset", "write_token_file", "DevicePrefetcher"):
        assert callable(getattr(data, name)), name
    assert "stall_timeout_s" in inspect.signature(watchdog.RankWatchdog.__init__).parameters
    assert callable(watchdog.RankWatchdog.tick) and watchdog.STALL_EXIT_CODE == 75
    assert {"eval_every", "watchdog_timeout_s", "resume"} <= set(inspect.signature(Trainer.__init__).parameters)
    assert callable(Callback.on_evaluate)
    for name in ("estimate_memory", "plan", "local_param_count", "main"):
        assert callable(getattr(planner, name)), name
    assert callable(BloomForCausalLM.to_hf) and callable(BloomForCausalLM.save_hf_pretrained) and callable(BloomConfig.to_hf)
    gen = inspect.signature(BloomForCausalLM.generate).parameters
    assert {"attention_mask", "do_sample", "temperature", "top_k", "top_p", "eos_token_id", "pad_token_id", "use_cache"} <= set(gen)


def test_diagnostics_entry_point_runs(capsys):
    """``python -m pipegoose_b200`` (bug-report helper): never raises, reports version and extension state."""
    import pipegoose_b200
    from pipegoose_b200.__main__ import main

    assert main() == 0
    out = capsys.readouterr().out
    assert pipegoose_b200.__version__ in out and "extension (_C.so)" in out and "distributed backends" in out
    import re
    import pathlib

    pyproject = (pathlib.Path(pipegoose_b200.__file__).resolve().parent.parent / "pyproject.toml").read_text()
    assert re.search(r'^version = "(.+)"$', pyproject, re.M).group(1) == pipegoose_b200.__version__
