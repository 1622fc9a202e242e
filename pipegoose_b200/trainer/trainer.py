"""A working trainer (the reference's Trainer.fit/train bodies are ``pass``, trainer/trainer.py:13-35): epochs over a
dataloader of ``{"input_ids", ["attention_mask"], ["labels"]}`` batches, loss/backward/step, callbacks, tokens/s
logging — plus what a long run needs: gradient accumulation (``no_sync`` on all but the last micro-step), global
gradient-norm clipping, an LR scheduler, periodic sharded checkpoints with resume, and a rank watchdog."""
from __future__ import annotations

import time
from typing import List, Optional

import torch
from torch import nn

from pipegoose_b200.trainer.callback import Callback
from pipegoose_b200.trainer.logger import DistributedLogger
from pipegoose_b200.trainer.state import TrainerStage, TrainerState, TrainerStatus


class Trainer:
    def __init__(self, module: nn.Module, train_loader, eval_loader=None, optim=None, num_epochs: int = 1,
                 callbacks: Optional[List[Callback]] = None, loggers: Optional[List[DistributedLogger]] = None,
                 parallel_context=None, log_every: int = 10, grad_accum_steps: int = 1,
                 max_grad_norm: Optional[float] = None, lr_scheduler=None, checkpoint_dir: Optional[str] = None,
                 checkpoint_every: int = 0, resume: bool = False, watchdog_timeout_s: Optional[float] = None,
                 max_steps: Optional[int] = None, keep_checkpoints: int = 2, moe_loss_weights=(0.01, 0.001),
                 eval_every: int = 0):
        """``grad_accum_steps``: micro-batches per optimizer step.  ``max_grad_norm``: clip the whole model's gradient
        norm (optim/clip.py); a step whose norm is not finite is skipped (gradients dropped, weights and optimizer state
        untouched, ``state.skipped_steps``).  ``lr_scheduler``: anything with ``step()`` (built on ``optim.optim`` for a
        ``DistributedOptimizer``).  ``checkpoint_dir`` + ``checkpoint_every``: sharded weights (nn.utils.save_pretrained)
        and optimizer / RNG / step state (save_training_state) every N optimizer steps, each in its own ``step_<n>``
        directory that becomes ``latest`` only when every rank has written its shard (``keep_checkpoints`` newest kept; one
        more is written when ``fit`` ends between two periodic ones); ``resume=True`` restores the
        latest one before training and skips the batches it had consumed.  ``max_steps``: stop at that optimizer step.  ``watchdog_timeout_s``: start a
        :class:`RankWatchdog` for the duration of ``fit``.  ``eval_every``: evaluate on ``eval_loader`` every N optimizer
        steps and once more when ``fit`` ends (``state.last_eval_loss``, the loggers' ``eval_loss``, ``on_evaluate``); the
        random generators are put back afterwards, so a run trains identically with and without evaluations."""
        assert grad_accum_steps >= 1
        self.module = module
        self.train_loader = train_loader
        self.eval_loader = eval_loader
        self.optim = optim
        self.num_epochs = num_epochs
        self.callbacks = sorted(callbacks or [], key=lambda c: c.order)
        self.loggers = loggers or []
        self.parallel_context = parallel_context
        self.log_every = log_every
        self.grad_accum_steps = grad_accum_steps
        self.max_grad_norm = max_grad_norm
        self.lr_scheduler = lr_scheduler
        self.checkpoint_dir = checkpoint_dir
        self.checkpoint_every = checkpoint_every
        self.resume = resume
        self.watchdog_timeout_s = watchdog_timeout_s
        self.moe_loss_weights = moe_loss_weights   # (load-balancing, router-z) weights for MoE models
        self.keep_checkpoints = max(1, keep_checkpoints)   # newest complete step directories kept on disk
        self.eval_every = int(eval_every or 0)
        self.max_steps = max_steps   # stop once this many optimizer steps exist in total (counting resumed ones)
        self.state = TrainerState()
        self._micro = 0
        self._skip_batches = 0
        self._resume = None               # set by load_checkpoint(): where in which epoch to continue, and the RNG states
        self._batches_in_epoch = 0
        self._epoch_start_rng = None      # torch CPU generator state when the current epoch's loader iterator was made

    def _device(self):
        return next(self.module.parameters()).device

    def _call(self, name, *args):
        import inspect

        for cb in self.callbacks:
            hook = getattr(cb, name)
            if name in ("on_fit_start", "on_fit_end") and not args:
                # the reference's signature is (trainer, pl_module); callbacks written as (trainer) keep working
                if len(inspect.signature(hook).parameters) >= 2:
                    hook(self, self.module)
                    continue
            hook(self, *args)

    def _log(self, msg):
        for lg in self.loggers:
            lg.info(msg)

    def _log_metrics(self, metrics: dict):
        for lg in self.loggers:
            if hasattr(lg, "log_metrics"):
                lg.log_metrics({k: v for k, v in metrics.items() if v == v})   # drop NaNs (e.g. no clipping configured)

    def _current_lr(self) -> float:
        groups = getattr(self.optim, "param_groups", None) or [{}]
        return float(groups[0].get("lr", float("nan")))

    def train_step(self, batch) -> torch.Tensor:
        """One micro-batch: forward + backward; the optimizer steps on every ``grad_accum_steps``-th call."""
        from contextlib import nullcontext

        dev = self._device()
        batch = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
        labels = batch.pop("labels", batch["input_ids"])
        first = self._micro == 0
        last = self._micro == self.grad_accum_steps - 1
        no_sync = getattr(self.module, "no_sync", None)
        # a pipelined module runs its backward INSIDE forward: the 1/k of gradient accumulation has to go in with the call
        pipelined = hasattr(self.module, "_pg_pipeline_engine") and self.grad_accum_steps > 1
        extra = {"loss_scale": 1.0 / self.grad_accum_steps} if pipelined else {}
        with (no_sync() if (no_sync is not None and not last) else nullcontext()):
            out = self.module(**batch, labels=labels, **extra)
            loss = out.loss if hasattr(out, "loss") else out[0]
            loss = self._add_router_losses(loss)
            if first:
                self.optim.zero_grad()   # after the forward, as in the reference's README loop
            (loss / self.grad_accum_steps if (self.grad_accum_steps > 1 and not pipelined) else loss).backward()
        self.state.tokens_seen += int(batch["input_ids"].numel())
        self._micro += 1
        if last:
            self._micro = 0
            if self.max_grad_norm is not None:
                from pipegoose_b200.optim.clip import clip_grad_norm_

                ctx = self.parallel_context if self.parallel_context is not None else _SingleProcess()
                self.state.last_grad_norm = float(clip_grad_norm_(self.optim, self.max_grad_norm, ctx))
                if self.state.last_grad_norm != self.state.last_grad_norm or self.state.last_grad_norm == float("inf"):
                    # a non-finite gradient (overflow, a poisoned batch): the clip coefficient would be NaN / 0 * inf and
                    # one step would write NaNs into the fp32 master weights for good.  The norm is a global quantity
                    # (all-reduced), so every rank takes this branch together: drop the gradients, keep the weights.
                    self._skip_step()
                    return loss
            self.optim.step()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            self.state.step += 1
            if self.checkpoint_dir and self.checkpoint_every and self.state.step % self.checkpoint_every == 0:
                self.save_checkpoint()
        return loss

    def _skip_step(self):
        inner = getattr(self.optim, "optim", self.optim)
        if hasattr(inner, "pending_grad_scale"):
            inner.pending_grad_scale = 1.0          # (what clip_grad_norm_ left for the fused Adam kernel)
        self.optim.zero_grad()
        self.state.skipped_steps += 1
        self._log(f"step {self.state.step + 1}: non-finite gradient norm, optimizer step skipped "
                  f"({self.state.skipped_steps} skipped so far)")

    def _add_router_losses(self, loss: torch.Tensor) -> torch.Tensor:
        """Mixture-of-experts layers push their load-balancing / router-z losses into the ExpertContext on every forward:
        add them with ``moe_loss_weights`` and drain the context (a pipelined module has done both inside its stages)."""
        from pipegoose_b200.nn.expert_parallel.expert_context import ExpertContext

        store = ExpertContext.get_instance()
        aux, z = store.pop_all_aux_loss(), store.pop_all_z_loss()
        w_aux, w_z = self.moe_loss_weights
        for weight, terms in ((w_aux, aux), (w_z, z)):
            terms = [t for t in terms if isinstance(t, torch.Tensor)]
            if weight and terms:
                loss = loss + weight * torch.stack([t.float().reshape(()) for t in terms]).sum().to(loss.dtype)
        return loss

    # ------------------------------------------------------------------ checkpoints
    # layout:  <checkpoint_dir>/step_00000500/{pytorch_model_tp_*_pp_*.bin, optimizer_tp_*_pp_*_dp_*.bin}
    #          <checkpoint_dir>/latest        <- name of the newest COMPLETE step directory (written last, atomically)
    _last_eval_step = -1
    _last_saved_step = -1

    def _evaluate_in_training(self):
        """A periodic evaluation that leaves no trace in the training run: eval mode and stage are restored by
        ``evaluate``; the random generators (a ``DataLoader`` iterator draws a base seed when it is created) here."""
        from pipegoose_b200.nn.utils import capture_rng_state, restore_rng_state

        rng = capture_rng_state()
        try:
            loss = float(self.evaluate())
        finally:
            restore_rng_state(rng)
        self.state.last_eval_loss = loss
        self._last_eval_step = self.state.step
        self._log(f"step {self.state.step} eval loss {loss:.4f}")
        self._log_metrics({"step": self.state.step, "eval_loss": loss})
        self._call("on_evaluate", loss)
        return loss

    def _replicas(self) -> int:
        from pipegoose_b200.distributed.parallel_mode import ParallelMode

        ctx = self.parallel_context
        return ctx.get_world_size(ParallelMode.DATA) if ctx is not None else 1

    def _tick(self):
        wd = getattr(self, "_watchdog", None)
        if wd is not None:
            wd.tick()

    def _barrier(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()

    def save_checkpoint(self):
        import os
        import shutil

        from pipegoose_b200.nn.utils import save_pretrained, save_training_state

        ctx = self.parallel_context
        assert ctx is not None, "checkpoints are sharded by (tp, pp, dp) rank: a ParallelContext is required"
        name = f"step_{self.state.step:08d}"
        path = os.path.join(self.checkpoint_dir, name)
        save_pretrained(self.module, ckp_path=path, parallel_context=ctx)
        # exact resume: the epoch's sampler permutation is drawn from the global torch RNG when the loader's iterator is
        # created, so the state the generator had THEN is what lets a resumed run re-create the same order
        extra = {"epoch": self.state.epoch, "tokens_seen": self.state.tokens_seen,
                 "batches_in_epoch": self._batches_in_epoch, "epoch_start_rng": self._epoch_start_rng,
                 "lr_scheduler": self.lr_scheduler.state_dict() if hasattr(self.lr_scheduler, "state_dict") else None}
        save_training_state(self.optim, path, ctx, step=self.state.step, extra=extra)
        self._barrier()                       # every rank's shard is on disk ...
        if ctx.get_global_rank() == 0:        # ... only then does the checkpoint become the one to resume from
            tmp = os.path.join(self.checkpoint_dir, f".latest.{os.getpid()}")
            with open(tmp, "w") as f:
                f.write(name)
            os.replace(tmp, os.path.join(self.checkpoint_dir, "latest"))
            done = sorted(d for d in os.listdir(self.checkpoint_dir) if d.startswith("step_"))
            for stale in done[:-self.keep_checkpoints]:
                shutil.rmtree(os.path.join(self.checkpoint_dir, stale), ignore_errors=True)
        self._barrier()
        self._tick()
        self._last_saved_step = self.state.step
        self._log(f"checkpoint written at step {self.state.step} -> {path}")

    def _latest_checkpoint(self) -> Optional[str]:
        import os

        marker = os.path.join(self.checkpoint_dir, "latest")
        if not os.path.exists(marker):
            return None
        with open(marker) as f:
            path = os.path.join(self.checkpoint_dir, f.read().strip())
        return path if os.path.isdir(path) else None

    def load_checkpoint(self) -> bool:
        """Restore the newest complete checkpoint of ``checkpoint_dir`` if there is one; returns whether it did."""
        from pipegoose_b200.nn.utils import from_pretrained, load_training_state

        ctx = self.parallel_context
        if ctx is None or not self.checkpoint_dir:
            return False
        path = self._latest_checkpoint()
        if path is None:
            return False
        from_pretrained(self.module, ckp_path=path, parallel_context=ctx)
        # RNG states are restored by train() AFTER the consumed batches were replayed (replaying draws from the
        # generators: the sampler permutation, the DataLoader's base seed)
        meta = load_training_state(self.optim, path, ctx, restore_rng=False)
        self.state.step = meta["step"]
        extra = meta.get("extra") or {}
        self.state.tokens_seen = extra.get("tokens_seen", 0)
        if self.lr_scheduler is not None and extra.get("lr_scheduler") is not None:
            self.lr_scheduler.load_state_dict(extra["lr_scheduler"])
        if extra.get("epoch_start_rng") is not None:
            self._resume = {"epoch": int(extra.get("epoch", 0)), "batches": int(extra.get("batches_in_epoch", 0)),
                            "epoch_start_rng": extra["epoch_start_rng"], "rng": meta.get("rng")}
        else:   # checkpoint of an older version: positional replay from the first epoch
            self._skip_batches = self.state.step * self.grad_accum_steps
            self._resume = {"epoch": 0, "batches": None, "epoch_start_rng": None, "rng": meta.get("rng")}
        self._log(f"resumed from step {self.state.step} ({path})")
        return True

    def fit(self):
        self.state.status = TrainerStatus.RUNNING
        if self.resume:
            self.load_checkpoint()
        watchdog = None
        if self.watchdog_timeout_s is not None and self.parallel_context is not None:
            from pipegoose_b200.utils.watchdog import RankWatchdog

            # dead peers are noticed through their heartbeat, a wedged main thread (this one, or a peer's, which leaves
            # this one blocked in a collective) through the progress ticks below
            watchdog = RankWatchdog(self.parallel_context, timeout_s=self.watchdog_timeout_s,
                                    stall_timeout_s=self.watchdog_timeout_s).start()
        self._watchdog = watchdog
        self._call("on_fit_start")
        try:
            self.train()
        finally:
            self._watchdog = None
            if watchdog is not None:
                watchdog.stop()
        if self.eval_every and self.eval_loader is not None and self._last_eval_step != self.state.step:
            self._evaluate_in_training()
        if (self.checkpoint_dir and self.checkpoint_every and self.state.step % self.checkpoint_every != 0
                and self.state.step != self._last_saved_step and self._micro == 0):
            self.save_checkpoint()      # periodic checkpoints are on: the run's last steps are not left unsaved
        self._call("on_fit_end")
        self.state.status = TrainerStatus.FINISHED
        return self.state

    def train(self):
        self.state.stage = TrainerStage.TRAINING
        self.module.train()
        t0, tok0 = time.time(), self.state.tokens_seen
        seen = 0
        resume, self._resume = self._resume, None
        first_epoch = resume["epoch"] if resume is not None else 0
        for epoch in range(first_epoch, self.num_epochs):
            self.state.epoch = epoch
            sampler = getattr(self.train_loader, "sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)     # DistributedSampler: another permutation every epoch, the same on resume
            self._call("on_epoch_start")
            skip_here = 0
            if resume is not None and epoch == first_epoch:
                if resume["epoch_start_rng"] is not None:
                    torch.set_rng_state(resume["epoch_start_rng"])   # same sampler permutation as the interrupted run
                    skip_here = resume["batches"]
                if not skip_here and not self._skip_batches:
                    self._restore_rng(resume)
                    resume = None
            self._epoch_start_rng = torch.get_rng_state()
            self._batches_in_epoch = 0
            for batch in self.train_loader:
                if self.max_steps is not None and self.state.step >= self.max_steps and self._micro == 0:
                    break
                seen += 1
                self._batches_in_epoch += 1
                if self._batches_in_epoch <= skip_here or seen <= self._skip_batches:
                    # consumed before the checkpoint this run resumed from
                    self._tick()
                    if resume is not None and (self._batches_in_epoch == skip_here or seen == self._skip_batches):
                        self._restore_rng(resume)   # from here on the run continues exactly where it was cut
                        resume = None
                    continue
                before = self.state.step
                loss = self.train_step(batch)
                self._tick()
                if self.state.step != before:
                    if self.state.step % self.log_every == 0:
                        self.state.last_loss = float(loss.item())
                        dt = max(time.time() - t0, 1e-9)
                        # this replica's tokens; the job's rate is that times the number of replicas
                        replicas = self._replicas()
                        rate = (self.state.tokens_seen - tok0) / dt
                        self._log(f"step {self.state.step} loss {self.state.last_loss:.4f} tokens/s {rate * replicas:.0f}"
                                  + (f" ({replicas} replicas x {rate:.0f})" if replicas > 1 else ""))
                        self._log_metrics({"step": self.state.step, "loss": self.state.last_loss,
                                           "tokens_per_s": rate, "job_tokens_per_s": rate * replicas,
                                           "tokens_seen": self.state.tokens_seen, "grad_norm": self.state.last_grad_norm,
                                           "lr": self._current_lr()})
                    self._call("on_step_end", loss)
                    if self.eval_every and self.eval_loader is not None and self.state.step % self.eval_every == 0:
                        self._evaluate_in_training()
            self._call("on_epoch_end")

    @staticmethod
    def _restore_rng(resume):
        from pipegoose_b200.nn.utils import restore_rng_state

        if resume.get("rng") is not None:
            restore_rng_state(resume["rng"])

    @torch.no_grad()
    def evaluate(self, across_replicas: bool = True) -> float:
        """Mean loss over the batches of ``eval_loader`` — of ALL data-parallel replicas' loaders (each replica evaluates
        its shard, sums and counts are combined over the DATA group: every rank returns the same number, and it does
        not depend on how many replicas shared the work).  That makes it a COLLECTIVE over the DATA group;
        ``across_replicas=False`` evaluates this replica's loader only (e.g. when one replica evaluates alone)."""
        assert self.eval_loader is not None
        # evaluate() may run in the middle of fit (from an on_step_end / on_epoch_end callback): the rest of training
        # must continue in the mode and stage it was in (dropout, router noise, training capacity factor)
        was_training, prev_stage = self.module.training, self.state.stage
        self.state.stage = TrainerStage.VALIDATING
        self.module.eval()
        try:
            dev = self._device()
            total, n = 0.0, 0
            for batch in self.eval_loader:
                batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
                labels = batch.pop("labels", batch["input_ids"])
                out = self.module(**batch, labels=labels)
                total += float((out.loss if hasattr(out, "loss") else out[0]).item())
                n += 1
                self._tick()
                self._add_router_losses(torch.zeros(()))   # evaluation reports the task loss; just drain the expert context
            ctx = self.parallel_context
            from pipegoose_b200.distributed.parallel_mode import ParallelMode

            if across_replicas and ctx is not None and ctx.get_world_size(ParallelMode.DATA) > 1:
                import torch.distributed as dist

                group = ctx.get_group(ParallelMode.DATA)
                acc = torch.tensor([total, float(n)], dtype=torch.float64,
                                   device=dev if dist.get_backend(group) == "nccl" else "cpu")
                dist.all_reduce(acc, group=group)
                total, n = float(acc[0]), int(acc[1])
            return total / max(n, 1)
        finally:
            self.module.train(was_training)
            self.state.stage = prev_stage


class _SingleProcess:
    """Stands in for a ParallelContext when the trainer runs without one (every group has one member)."""

    def get_world_size(self, mode):
        return 1
