"""Checkpoint I/O (parity: reference nn/utils.py:11-50): one file per (tp_rank, pp_rank) named
``pytorch_model_tp_{tp}_pp_{pp}.bin``.  Only data-parallel rank 0 writes (the reference has every
replica write the same path) and the directory is created when missing.  Next to each shard goes
``<shard>.layout.json`` (which keys were cut along which dimension), which ``nn/checkpoint_convert.py`` uses to
merge the shards into one state dict or to re-cut them for another tensor-parallel size, offline."""
from __future__ import annotations

import os
from pathlib import Path

import torch
from torch import nn

from pipegoose_b200.constants import CHECKPOINT_PATH_NAME, CHECKPOINT_WEIGHTS_NAME
from pipegoose_b200.distributed.parallel_context import ParallelContext
from pipegoose_b200.distributed.parallel_mode import ParallelMode


def _ckpt_file(ckp_path: str, ckp_name: str, parallel_context: ParallelContext) -> str:
    tp_rank = parallel_context.get_local_rank(ParallelMode.TENSOR)
    pp_rank = parallel_context.get_local_rank(ParallelMode.PIPELINE)
    return os.path.join(ckp_path, ckp_name.format(tp_rank, pp_rank))


def _atomic_save(obj, path: str):
    """Write next to the target and rename: a job killed mid-write never leaves a truncated checkpoint behind (the
    previous complete file stays in place until the new one is whole)."""
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def from_pretrained(module: nn.Module, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                    ckp_name: str = CHECKPOINT_WEIGHTS_NAME, strict: bool = True):
    """Load this rank's (tp, pp) shard into an already parallelized ``module``.  ``strict`` (default, as the
    reference's ``load_state_dict``): every parameter and buffer of the module must be in the checkpoint."""
    path = _ckpt_file(ckp_path, ckp_name, parallel_context)
    if not os.path.exists(path):
        raise ValueError(f"ckp_path {path} does not exist")
    state_dict = torch.load(path, map_location="cpu")
    with torch.no_grad():
        own = module.state_dict()
        # strict, like the reference's ``module.load_state_dict``: a stale or partial shard (another model config,
        # parallel layout or MoE mapping) must not resume silently with randomly initialised leftovers
        unexpected = [k for k in state_dict if k not in own]
        if unexpected:
            raise KeyError(f"unexpected key(s) {unexpected[:5]} in checkpoint {path}")
        missing = [k for k in own if k not in state_dict]
        if missing and strict:
            raise KeyError(f"checkpoint {path} lacks {len(missing)} parameter(s) / buffer(s) of this module, e.g. "
                           f"{missing[:5]}")
        for k, v in state_dict.items():
            if tuple(own[k].shape) != tuple(v.shape):
                raise ValueError(f"shape mismatch for {k} in {path}: checkpoint {tuple(v.shape)}, module "
                                 f"{tuple(own[k].shape)} (different model config or parallel layout?)")
        for k, v in state_dict.items():
            own[k].copy_(v)
    return module


def save_pretrained(module: nn.Module, ckp_name: str = CHECKPOINT_WEIGHTS_NAME, ckp_path: str = CHECKPOINT_PATH_NAME,
                    parallel_context: ParallelContext = None):
    """Save this rank's (tp, pp) shard; data-parallel replicas > 0 skip the write."""
    Path(ckp_path).mkdir(parents=True, exist_ok=True)
    if parallel_context.get_local_rank(ParallelMode.DATA) == 0:
        state = {k: v.detach().cpu() for k, v in module.state_dict().items()}
        path = _ckpt_file(ckp_path, ckp_name, parallel_context)
        _atomic_save(state, path)
        # how every key was cut, for the offline tools (nn/checkpoint_convert.py: consolidate / reshard)
        from pipegoose_b200.nn.checkpoint_convert import write_layout

        write_layout(module, path, parallel_context)
    if parallel_context.get_world_size(ParallelMode.DATA) > 1:
        import torch.distributed as dist

        dist.barrier(group=parallel_context.get_group(ParallelMode.DATA))


# ------------------------------------------------------------------------------------------------
# resume: optimizer shards + RNG + step (the reference only saves model weights)
# ------------------------------------------------------------------------------------------------
def _optim_file(ckp_path: str, parallel_context: ParallelContext) -> str:
    from pipegoose_b200.constants import CHECKPOINT_OPTIM_NAME

    return os.path.join(ckp_path, CHECKPOINT_OPTIM_NAME.format(
        parallel_context.get_local_rank(ParallelMode.TENSOR), parallel_context.get_local_rank(ParallelMode.PIPELINE),
        parallel_context.get_local_rank(ParallelMode.DATA)))


def _layout(parallel_context: ParallelContext) -> dict:
    return {"tp": parallel_context.tensor_parallel_size, "pp": parallel_context.pipeline_parallel_size,
            "dp": parallel_context.data_parallel_size}


def save_training_state(optim, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                        step: int = 0, extra: dict = None):
    """Every rank writes ITS optimizer shard (ZeRO-1: 1/dp of the fp32 master weights and moments) as
    ``optimizer_tp_{tp}_pp_{pp}_dp_{dp}.bin`` together with the step counter, the RNG states and the parallel
    layout the shard belongs to."""
    Path(ckp_path).mkdir(parents=True, exist_ok=True)
    sd = optim.state_dict()

    def to_cpu(x):
        if isinstance(x, torch.Tensor):
            return x.detach().cpu()
        if isinstance(x, dict):
            return {k: to_cpu(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(to_cpu(v) for v in x)
        return x

    blob = {"optimizer": to_cpu(sd), "step": int(step), "layout": _layout(parallel_context),
            "rng": capture_rng_state(), "extra": extra or {}}
    _atomic_save(blob, _optim_file(ckp_path, parallel_context))


def capture_rng_state() -> dict:
    """Every generator a training loop may draw from: torch CPU + CUDA, Python ``random``, numpy."""
    import random

    state = {"torch": torch.get_rng_state(),
             "cuda": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
             "python": random.getstate()}
    try:
        import numpy as np

        state["numpy"] = np.random.get_state()
    except Exception:  # pragma: no cover - numpy missing
        state["numpy"] = None
    return state


def restore_rng_state(state: dict):
    import random

    torch.set_rng_state(state["torch"])
    if state.get("cuda") is not None and torch.cuda.is_available():
        torch.cuda.set_rng_state(state["cuda"])
    if state.get("python") is not None:
        random.setstate(state["python"])
    if state.get("numpy") is not None:
        import numpy as np

        np.random.set_state(state["numpy"])


def load_training_state(optim, ckp_path: str = CHECKPOINT_PATH_NAME, parallel_context: ParallelContext = None,
                        restore_rng: bool = True) -> dict:
    """Restore what :func:`save_training_state` wrote; returns ``{"step": ..., "extra": ...}`` (plus ``"rng"``, the saved
    generator states, when ``restore_rng=False`` leaves restoring them to the caller)."""
    path = _optim_file(ckp_path, parallel_context)
    if not os.path.exists(path):
        raise ValueError(f"optimizer checkpoint {path} does not exist")
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if blob["layout"] != _layout(parallel_context):
        raise ValueError(f"checkpoint was written for layout {blob['layout']}, this job runs {_layout(parallel_context)}: "
                         "optimizer shards are tied to the parallel layout")
    optim.load_state_dict(blob["optimizer"])
    meta = {"step": blob["step"], "extra": blob["extra"]}
    if restore_rng:
        restore_rng_state(blob["rng"])
    else:
        meta["rng"] = blob["rng"]   # the caller restores it (the Trainer: after it replayed the consumed batches)
    return meta
